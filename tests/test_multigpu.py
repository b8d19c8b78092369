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


@pytest.mark.parametrize("arch,transport,port", [("dsv2", "fused", 29581), ("llama", "fused", 29582), ("dsv2", "nccl", 29583)])
def test_serving_chain_parity(arch, transport, port):
    """The product path: LLMEngine -> ChainPipeline (shared-memory launch ring, fused P2P hand-off inside per-stage CUDA graphs),
    greedy + seeded sampled requests, against a single-GPU engine."""
    out = _torchrun("serving_chain_parity.py", [arch, transport], port=port)
    assert "SERVING_OK" in out, out[-3000:]


@pytest.mark.parametrize("version,port", [("v2", 29574), ("v1", 29578)])
def test_expert_parallel_moe(version, port):
    """v2: sender-side slot reservation + arrival wait fused into the grouped GEMM; v1: regroup kernels on the receive side."""
    out = _torchrun("ep_parity.py", [version], port=port)
    assert "EP_OK " + version in out, out[-3000:]


@pytest.mark.parametrize("version,port", [("v2", 29584), ("v1", 29585)])
def test_expert_parallel_stress(version, port):
    """Several EP layers sharing one buffer set, called back to back with no host synchronisation, token counts changing from step
    to step (prefill chunk <-> decode batch), two geometries; every output must equal the local MoE bit for bit."""
    out = _torchrun("ep_stress.py", [version], port=port)
    assert "EP_STRESS_OK " + version in out, out[-3000:]


def test_expert_parallel_whole_model():
    """Data-parallel attention + expert-parallel MoE for every layer, graph-captured decode loop (bench.py --parallelism ep)."""
    out = _torchrun("ep_model_parity.py", [], port=29576)
    assert "EP_MODEL_OK" in out, out[-3000:]


def test_generate_cli_expert_parallel(tmp_path):
    """``torchrun generate.py --expert_parallel`` (experts sharded at load, fused all-to-all) prints the same greedy text as a
    single-GPU run of the whole model."""
    import torch

    sys.path.insert(0, HERE)
    from helpers import GPU_DSV2
    from mlx_sharding_b200.utils.checkpoint import write_synthetic_checkpoint

    ckpt = write_synthetic_checkpoint(str(tmp_path / "dsv2"), GPU_DSV2, dtype=torch.bfloat16, seed=11)
    root = os.path.dirname(HERE)
    common = ["--model", ckpt, "--prompt", "expert parallel", "--max_tokens", "12", "--no_chat_template"]
    one = subprocess.run([sys.executable, os.path.join(root, "generate.py"), *common], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES="0"))
    assert one.returncode == 0, one.stderr[-2000:]
    ep = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                         "--master-port", "29577", os.path.join(root, "generate.py"), *common, "--expert_parallel"],
                        capture_output=True, text=True, timeout=420)
    assert ep.returncode == 0, (ep.stdout + ep.stderr)[-3000:]
    # NCCL prints a version banner on stdout when NCCL_DEBUG=VERSION is set in the environment: not part of the generation
    text = lambda out: "".join(l for l in out.split("==========")[0].splitlines(True) if not l.startswith("NCCL version"))
    assert "Generation:" in ep.stdout
    assert text(one.stdout) == text(ep.stdout) and len(text(one.stdout)) > 0
