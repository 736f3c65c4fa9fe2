"""Data-parallel train step on real GPUs: needs at least two devices, skipped otherwise.

Runs scripts/dp_parity.py under torchrun (one process per GPU, rendezvous on 127.0.0.1) for the three gradient exchanges:
the copy-engine exchange over NVLink peer memory with the decoder span's Adam update sharded over the ranks (default:
reduce-scatter, each rank updates its chunk, all-gather of the updated weights), the same exchange as a plain all-reduce
(PCAA_DP_SHARD_ADAM=0) and the NCCL all-reduce.  The script checks
that the exchanged gradient equals the sum of single-process per-shard gradients, that the weights are the Adam update of
the mean gradient, and that all replicas are bit-identical after the step."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["peer", "peer-allreduce", "nccl"])
def test_dp_parity_two_ranks(exchange):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, PCAA_DP_EXCHANGE=exchange.split("-")[0], PCAA_DP_SHARD_ADAM="0" if exchange == "peer-allreduce" else "1")
    port = 29600 + (os.getpid() % 300) + ["peer", "peer-allreduce", "nccl"].index(exchange)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "scripts", "dp_parity.py")],
                       capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    print("\n".join(l for l in r.stdout.splitlines() if l.startswith("dp_parity")))
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("dp_parity world")][-1]
    assert line.endswith("-> OK"), line
    assert ("peer copies" in line) == exchange.startswith("peer"), line
    assert ("sharded decoder update" in line) == (exchange == "peer"), line


@pytest.mark.gpu
def test_syncbn_two_ranks_equal_one_process_on_the_global_batch():
    """PCAATrainer(sync_bn=True): BatchNorm statistics all-reduced over the ranks -> the 2-rank iteration is the
    single-process iteration of the concatenated batch (gradients, running statistics, losses)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29900 + (os.getpid() % 90)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "scripts", "dp_parity.py"), "--sync-bn"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    print("\n".join(l for l in r.stdout.splitlines() if l.startswith("dp_parity")))
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("dp_parity syncbn")][-1]
    assert line.endswith("-> OK"), line
