"""Multi-GPU tests (torchrun, one rank per GPU): fused-P2P / NCCL pipeline parity with a single-GPU run, and
expert-parallel MoE parity.  Skipped unless >= 2 CUDA devices are visible."""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
HERE = os.path.dirname(os.path.abspath(__file__))


def _torchrun(script, args, nproc=2, port=29571, timeout=420):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "mgpu", script), *args]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    return r.stdout + r.stderr


@pytest.mark.parametrize("arch,transport,port", [("dsv2", "fused", 29571), ("llama", "fused", 29572), ("dsv2", "nccl", 29573),
                                                 ("dsv2", "fused,half", 29575)])
def test_pipeline_parity(arch, transport, port):
    """``fused,half``: the stage boundary sits between the attention and the MLP block (o-proj epilogue does the P2P store)."""
    out = _torchrun("pipeline_parity.py", [arch] + transport.split(","), port=port)
    assert "PARITY_OK" in out, out[-3000:]


def test_expert_parallel_moe():
    out = _torchrun("ep_parity.py", [], port=29574)
    assert "EP_OK" in out, out[-3000:]


def test_expert_parallel_whole_model():
    """Data-parallel attention + expert-parallel MoE for every layer, graph-captured decode loop (bench.py --parallelism ep)."""
    out = _torchrun("ep_model_parity.py", [], port=29576)
    assert "EP_MODEL_OK" in out, out[-3000:]
