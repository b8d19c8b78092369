#!/usr/bin/env python
"""HTTP-level serving benchmark (BASELINE config 4: Llama-3-8B bf16, 4 pipeline stages, OpenAI ``/v1/chat/completions``).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 bench/http_bench.py \
        --model synthetic:llama3-8b --concurrency 256 --max-tokens 64

Rank 0 runs the product's HTTP server (``server/openai_api.py``: ``ModelProvider`` -> ``LLMEngine`` -> ``ChainPipeline`` with the
shared-memory launch ring and the fused P2P hand-off) in a thread; the other ranks are ordinary stage workers
(``shard_server.serve_chain``).  ``--concurrency`` client threads then POST streaming chat requests with the API's DEFAULT sampling
parameters (temperature 1.0, top_p 1.0 — the reference's defaults, shard/openai_api.py:206-215), i.e. sampled requests, which replay
the same CUDA graphs as greedy ones because the sampling parameters travel in the step block.  Reports completion tokens/s over
the whole run (wall clock, host side — this is an end-to-end number through TCP, JSON and SSE), TTFT percentiles and the engine's
graph-replay counters.  The load generator is a separate process (``--client``) so its threads do not contend for the server's GIL.
Two untimed warm-up rounds precede the measured one (graph capture of every batch bucket, allocator warm-up)."""
import argparse
import http.client
import json
import os
import statistics
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def client(port, prompt, max_tokens, out, i, extra):
    body = dict(messages=[{"role": "user", "content": prompt}], stream=True, max_tokens=max_tokens, **extra)
    t0 = time.perf_counter()
    c = http.client.HTTPConnection("127.0.0.1", port, timeout=600)
    c.request("POST", "/v1/chat/completions", json.dumps(body), {"Content-Type": "application/json"})
    r = c.getresponse()
    first, n = None, 0
    buf = b""
    while True:
        chunk = r.read1(65536) if hasattr(r, "read1") else r.read(4096)
        if not chunk:
            break
        buf += chunk
        while b"\n\n" in buf:
            ev, buf = buf.split(b"\n\n", 1)
            if not ev.startswith(b"data: "):
                continue
            data = ev[6:]
            if data == b"[DONE]":
                continue
            if first is None:
                first = time.perf_counter() - t0
            n += 1
    c.close()
    out[i] = (first, n, time.perf_counter() - t0, r.status)


def run_round(port, conc, max_tokens, extra, offset=0):
    """One client process: ``conc`` threads, one streaming request each; returns the raw per-request records."""
    out = [None] * conc
    prompts = ["Request %03d: write a short note about pipeline parallel inference on NVLink-connected GPUs. " % (offset + i) for i in range(conc)]
    th = [threading.Thread(target=client, args=(port, prompts[i], max_tokens, out, i, extra)) for i in range(conc)]
    t0 = time.time()
    for t in th:
        t.start()
    for t in th:
        t.join()
    return dict(t0=t0, t1=time.time(), records=[o for o in out if o])


def summarise(parts, conc):
    recs = [r for p in parts for r in p["records"]]
    wall = max(p["t1"] for p in parts) - min(p["t0"] for p in parts)
    ok = [o for o in recs if o[3] == 200 and o[0] is not None]
    toks = sum(o[1] for o in ok)
    ttft = sorted(o[0] for o in ok)
    pct = lambda q: ttft[min(len(ttft) - 1, int(q * len(ttft)))] if ttft else None
    return dict(requests=conc, ok=len(ok), sse_chunks=toks, wall_s=round(wall, 3), sse_chunks_per_s=round(toks / wall, 1),
                ttft_p50_ms=round(pct(0.5) * 1e3, 1) if ttft else None, ttft_p90_ms=round(pct(0.9) * 1e3, 1) if ttft else None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="synthetic:llama3-8b")
    ap.add_argument("--concurrency", type=int, default=256)
    ap.add_argument("--max-tokens", type=int, default=64)
    ap.add_argument("--max-batch", type=int, default=64)
    ap.add_argument("--port", type=int, default=18080)
    ap.add_argument("--greedy", action="store_true", help="temperature 0 instead of the API defaults")
    ap.add_argument("--api-workers", type=int, default=8, help="front-end worker processes of the server (0 = in-process HTTP threads)")
    ap.add_argument("--client-procs", type=int, default=4, help="load-generator processes (the concurrency is split between them)")
    ap.add_argument("--offset", type=int, default=0, help="(internal) first request number of this load-generator process")
    ap.add_argument("--client", action="store_true", help="(internal) load-generator process: run one round against --port, print JSON")
    a = ap.parse_args()
    if a.client:
        print(json.dumps(run_round(a.port, a.concurrency, a.max_tokens, dict(temperature=0.0) if a.greedy else {}, a.offset)), flush=True)
        return

    import torch

    from mlx_sharding_b200.server import openai_api as api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    argv = ["--model", a.model, "--port", str(a.port), "--max-batch", str(a.max_batch), "--kv-pages", "4096", "--log-level", "WARNING"]
    if rank != 0:
        return api.main(argv)                       # stage worker: returns when rank 0 shuts the chain down
    # rank 0: the same start-up as ``mlx-sharding-api`` (layer split, ModelProvider, HTTP server), but in a thread
    args = api.build_arg_parser().parse_args(argv)
    if world > 1:
        from mlx_sharding_b200.config import ModelConfig
        from mlx_sharding_b200.parallel.partition import balanced_split
        from mlx_sharding_b200.parallel.transport import init_distributed
        from mlx_sharding_b200.utils.checkpoint import get_model_path

        init_distributed(device=args.device)
        cfg = ModelConfig.from_path(get_model_path(args.model))
        spec = balanced_split(cfg, world)[0]
        args.start_layer, args.end_layer = spec.start_layer, spec.end_layer
    args.static_dir = os.path.join(os.path.dirname(os.path.abspath(api.__file__)), "static")
    provider = api.ModelProvider(args, [])
    th = threading.Thread(target=api.run, args=(args.host, args.port, provider, args.static_dir), kwargs=dict(api_workers=a.api_workers),
                          daemon=True)
    th.start()
    for _ in range(600):
        try:
            c = http.client.HTTPConnection("127.0.0.1", a.port, timeout=2)
            c.request("GET", "/health")
            if c.getresponse().status == 200:
                break
        except OSError:
            time.sleep(0.2)
    # the load generator is a separate process (its threads must not share the server's GIL)
    import subprocess

    def round_(max_tokens):
        P = max(1, min(a.client_procs, a.concurrency))
        per = [a.concurrency // P + (1 if i < a.concurrency % P else 0) for i in range(P)]
        procs, off = [], 0
        for n in per:
            cmd = [sys.executable, os.path.abspath(__file__), "--client", "--port", str(a.port), "--concurrency", str(n),
                   "--max-tokens", str(max_tokens), "--offset", str(off)] + (["--greedy"] if a.greedy else [])
            procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=dict(os.environ, CUDA_VISIBLE_DEVICES="")))
            off += n
        parts = []
        for p in procs:
            so, se = p.communicate(timeout=1200)
            assert p.returncode == 0, se[-2000:]
            parts.append(json.loads(so.strip().splitlines()[-1]))
        return summarise(parts, a.concurrency)

    warm = round_(min(a.max_tokens, 16))
    warm2 = round_(a.max_tokens)        # second warm-up at the measured length: every decode bucket / context bucket is captured
    eng = provider.engine
    st0 = dict(eng.stats)
    # sample the engine's token counter during the measured round: the steady-state rate is the slope over the middle half of the
    # tokens (a burst of --concurrency simultaneous connections spends its first part in connection set-up, chat templating and
    # prefill, its last part draining stragglers — neither is the decode rate the device-side bench reports)
    samples, stop_mon = [], threading.Event()

    def monitor():
        while not stop_mon.is_set():
            samples.append((time.perf_counter(), eng.stats["decode_tokens"]))
            time.sleep(0.02)

    mon = threading.Thread(target=monitor, daemon=True)
    mon.start()
    res = round_(a.max_tokens)
    stop_mon.set()
    mon.join()
    total = eng.stats["decode_tokens"] - st0["decode_tokens"]
    lo = next((s_ for s_ in samples if s_[1] - st0["decode_tokens"] >= 0.25 * total), None)
    hi = next((s_ for s_ in samples if s_[1] - st0["decode_tokens"] >= 0.75 * total), None)
    steady = round((hi[1] - lo[1]) / (hi[0] - lo[0]), 1) if lo and hi and hi[0] > lo[0] else None
    pipe = eng.pipe
    gc = getattr(pipe, "gcache", None)
    out = {"bench": "HTTP /v1/chat/completions, streaming, " + ("temperature 0" if a.greedy else "API default sampling (temperature 1.0)"),
           "model": a.model, "n_gpus": world, "api_workers": a.api_workers, "concurrency": a.concurrency, "max_tokens": a.max_tokens, "pipeline": type(pipe).__name__,
           "hand_off": getattr(getattr(pipe, "plane", None), "name", "local"),
           "control": type(getattr(pipe, "ctl", None)).__name__ if hasattr(pipe, "ctl") else None,
           "warmup_round": warm, "warmup_round_2": warm2, "measured_round": res,
           "engine_decode_tokens": eng.stats["decode_tokens"] - st0["decode_tokens"], "engine_steps": eng.stats["steps"] - st0["steps"],
           "steady_decode_tokens_per_s": steady,
           "steady_note": "slope of the engine's decode-token counter between 25 % and 75 % of the round's tokens (all streams active)",
           "tokens_per_s": round((eng.stats["decode_tokens"] - st0["decode_tokens"] + res["ok"]) / res["wall_s"], 1),
           "graph_replays_stage0": gc.replays if gc is not None else None, "graph_captures_stage0": gc.captures if gc is not None else None}
    print(json.dumps(out), flush=True)
    eng.shutdown()
    if hasattr(pipe, "shutdown"):
        pipe.shutdown()
    os._exit(0)     # the HTTP server thread and the engine thread are daemons; workers have left their loop


if __name__ == "__main__":
    main()
