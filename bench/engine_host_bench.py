#!/usr/bin/env python
"""Host-side cost of the serving loop, with the GPUs taken out: ``LLMEngine`` drives a stub pipeline that packs the step block exactly like
``ChainPipeline.submit`` (``pack_step``: tag | sampling block | BatchMeta | tokens) and answers with random tokens immediately.

    python bench/engine_host_bench.py [--groups 4 --batch 64 --requests 512 --max-tokens 64]

The printed ``us_per_step`` is the Python time the scheduler + step packing + result handling need per micro-batch step, i.e. the
floor of the inter-step interval one engine process can sustain; ``tokens_per_s_ceiling`` = batch / that.  (Runs on CPU; no GPU.)"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--groups", type=int, default=4)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--requests", type=int, default=512)
    ap.add_argument("--prompt", type=int, default=24)
    ap.add_argument("--max-tokens", type=int, default=64)
    ap.add_argument("--temperature", type=float, default=1.0)
    a = ap.parse_args()

    import torch

    from mlx_sharding_b200.engine.core import LLMEngine, StepOutput
    from mlx_sharding_b200.engine.sampler import SamplingParams
    from mlx_sharding_b200.parallel.graph_decode import pack_step

    class StubPipeline:
        num_stages = a.groups

        def __init__(self):
            self.seq, self.pack_s = 0, 0.0

        def submit(self, inp):
            self.seq += 1
            t0 = time.perf_counter()
            pack_step(self.seq, inp.meta, inp.tokens, inp.params, inp.contexts, inp.rng, inp.is_prefill, pad_decode=True)
            self.pack_s += time.perf_counter() - t0
            n = inp.meta.num_seqs
            return StepOutput(tokens=torch.randint(3, 1000, (n,)).tolist(), logprobs=[0.0] * n)

        def wait(self, h):
            return h

        def reset(self):
            pass

    pipe = StubPipeline()
    eng = LLMEngine(pipe, 1 << 15, 64, num_groups=a.groups, max_seqs_per_group=a.batch, max_prefill_tokens=2048)
    prm = SamplingParams(temperature=a.temperature)
    reqs = [eng.submit(list(range(3, 3 + a.prompt)), prm, max_tokens=a.max_tokens, eos_token_id=None) for _ in range(a.requests)]
    t0 = time.perf_counter()
    eng.drain()
    dt = time.perf_counter() - t0
    steps, toks = eng.stats["steps"], eng.stats["decode_tokens"]
    print(json.dumps({"bench": "engine host loop, stub pipeline", "groups": a.groups, "batch": a.batch, "requests": a.requests,
                      "steps": steps, "decode_tokens": toks, "wall_s": round(dt, 3), "us_per_step": round(dt / steps * 1e6, 1),
                      "us_per_step_pack": round(pipe.pack_s / steps * 1e6, 1),
                      "tokens_per_s_ceiling": round(toks / dt, 0)}))
    assert all(r.finished for r in reqs)


if __name__ == "__main__":
    main()
