"""Engine / scheduler tests on the CPU reference path (tiny Llama)."""
import torch

from helpers import TINY_DSV2, TINY_LLAMA, run_sequence
from mlx_sharding_b200.config import ModelConfig
from mlx_sharding_b200.engine.core import LLMEngine, stopping_criteria
from mlx_sharding_b200.engine.sampler import SamplingParams
from mlx_sharding_b200.models import build_stage
from mlx_sharding_b200.parallel.pipeline import LocalPipeline
from mlx_sharding_b200.utils.checkpoint import random_state_dict


def _engine(C=TINY_LLAMA, ranges=None, **kw):
    cfg = ModelConfig.from_dict(C)
    sd = dict(random_state_dict(cfg, dtype=torch.float32))
    ranges = ranges or [(0, cfg.num_hidden_layers)]
    models = [build_stage(cfg, cfg.shard(s, e), torch.float32).load_state(sd) for s, e in ranges]
    pipe = LocalPipeline.from_models(models, num_pages=64, page_size=16)
    return models, LLMEngine(pipe, num_pages=64, page_size=16, **kw)


def _greedy_oracle(models, prompt, n):
    outs = run_sequence(models, prompt, n - 1)
    return [int(o.argmax()) for o in outs]


def test_greedy_matches_single_sequence_oracle():
    models, eng = _engine()
    prompt = [3, 9, 27, 81, 243, 11]
    assert eng.generate(prompt, SamplingParams(temperature=0.0), max_tokens=6) == _greedy_oracle(models, prompt, 6)


def test_batch_invariance_and_chunked_prefill():
    """n concurrent sequences through the micro-batch scheduler == n sequential runs."""
    models, eng = _engine(TINY_DSV2, ranges=[(0, 2), (2, 4)], num_groups=2, max_prefill_tokens=5)
    prompts = [[5, 6, 7, 8, 9, 10, 11], [100, 50], [1, 2, 3], [42] * 9, [7, 300, 12, 13]]
    reqs = [eng.submit(p, SamplingParams(), max_tokens=5) for p in prompts]
    eng.drain()
    for p, r in zip(prompts, reqs):
        assert r.finished and r.finish_reason == "length"
        assert r.output == _greedy_oracle(models, p, 5), p
    assert eng.table.alloc.num_free == 63  # every page returned


def test_stop_conditions():
    models, eng = _engine()
    prompt = [3, 9, 27]
    ref = _greedy_oracle(models, prompt, 8)
    k = next(i for i in range(1, 8) if ref[i] not in ref[:i])
    r = eng.submit(prompt, SamplingParams(), max_tokens=8, eos_token_id=ref[k])
    eng.drain()
    assert r.finish_reason == "stop" and r.output == ref[:k + 1]
    j = next(i for i in range(2, 8) if all(ref[m:m + 2] != ref[i - 1:i + 1] for m in range(i - 1)))
    r = eng.submit(prompt, SamplingParams(), max_tokens=8, stop_id_sequences=[ref[j - 1:j + 1]])
    eng.drain()
    assert r.finish_reason == "stop" and r.output == ref[:j + 1]
    assert stopping_criteria([1, 2, 3], [[2, 3]], None) == (True, 2)
    assert stopping_criteria([1, 2, 3], [[9]], 3) == (True, 1)
    assert stopping_criteria([1, 2, 3], [], None) == (False, 0)


def test_threaded_engine_streams_events():
    models, eng = _engine()
    eng.start()
    try:
        r = eng.submit([4, 5, 6], SamplingParams(logprobs=3), max_tokens=4)
        evs = list(r)
        assert len(evs) == 4 and evs[-1].finished and evs[-1].finish_reason == "length"
        assert all(len(e.top) == 3 for e in evs)
        assert all(e.logprob <= 0 for e in evs)
        # the chosen greedy token is the top-1 logprob entry
        assert all(abs(max(e.top.values()) - e.logprob) < 1e-5 for e in evs)
        assert r.ttft is not None and r.ttft > 0
    finally:
        eng.shutdown()


def test_sampling_params_affect_output():
    models, eng = _engine()
    prompt = [3, 9, 27]
    greedy = eng.generate(prompt, SamplingParams(), max_tokens=6)
    biased = eng.generate(prompt, SamplingParams(logit_bias={7: 100.0}), max_tokens=3)
    assert biased == [7, 7, 7]
    pen = eng.generate(prompt, SamplingParams(logit_bias={7: 5.0}, repetition_penalty=50.0,
                                              repetition_context_size=20), max_tokens=6)
    assert pen.count(7) <= 1 or pen != [7] * 6
    torch.manual_seed(0)
    hot = eng.generate(prompt, SamplingParams(temperature=5.0, top_p=0.95, seed=1), max_tokens=12)
    assert hot != greedy[:12] or True  # sampling path executes; distribution checked in test_sampler
    assert len(hot) == 12


def test_oversized_request_rejected():
    models, eng = _engine()
    r = eng.submit(list(range(1, 200)), SamplingParams(), max_tokens=2000)
    eng.drain()
    assert isinstance(r.error, MemoryError)


def _prefix_engine(num_pages=64, page_size=4, ranges=None, **kw):
    cfg = ModelConfig.from_dict(TINY_LLAMA)
    sd = dict(random_state_dict(cfg, dtype=torch.float32))
    ranges = ranges or [(0, cfg.num_hidden_layers)]
    models = [build_stage(cfg, cfg.shard(s, e), torch.float32).load_state(sd) for s, e in ranges]
    pipe = LocalPipeline.from_models(models, num_pages=num_pages, page_size=page_size)
    return models, LLMEngine(pipe, num_pages=num_pages, page_size=page_size, prefix_cache=True, **kw)


def test_prefix_cache_reuses_prompt_pages_and_keeps_outputs():
    """Automatic prefix caching: a second prompt sharing a prefix skips the cached full pages (less prefill work) and still
    produces exactly the tokens of an uncached run; works across a 2-stage pipeline (same page ids on every stage)."""
    models, eng = _prefix_engine(ranges=[(0, 2), (2, 4)])
    system = [7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17]                     # 11 tokens: 2 full pages of 4 + 3
    a, b = system + [100, 101], system + [200]
    g = SamplingParams(temperature=0.0)
    assert eng.generate(a, g, max_tokens=5) == _greedy_oracle(models, a, 5)
    assert eng.stats["prefix_cached_tokens"] == 0
    before = eng.stats["prefill_tokens"]
    assert eng.generate(b, g, max_tokens=5) == _greedy_oracle(models, b, 5)
    assert eng.stats["prefix_cached_tokens"] == 8                          # two full pages re-used
    assert eng.stats["prefill_tokens"] - before == len(b) - 8
    # the very same prompt again: everything but the page holding the last prompt token is re-used
    assert eng.generate(a, g, max_tokens=5) == _greedy_oracle(models, a, 5)
    assert eng.stats["prefix_cached_tokens"] == 8 + 12
    # concurrent requests on shared pages + a batch of unrelated ones
    reqs = [eng.submit(p, g, max_tokens=4) for p in (a, b, [5, 5, 5, 5, 5, 5], system + [300, 301, 302])]
    eng.drain()
    for r, p in zip(reqs, (a, b, [5, 5, 5, 5, 5, 5], system + [300, 301, 302])):
        assert r.output == _greedy_oracle(models, p, 4)
    # bookkeeping: nothing is referenced any more, every page is either free or evictable; eviction returns all of them
    cache = eng.table.prefix
    assert all(v == 0 for v in cache.refs.values()) and len(cache.lru) == len(cache.refs) > 0
    assert eng.table.alloc.num_free == 64 - 1
    cache.evict(10 ** 6)
    assert len(eng.table.alloc._free) == 64 - 1 and not cache.by_digest


def test_prefix_cache_evicts_under_memory_pressure():
    """A pool too small to keep old prefixes: cached pages are evicted LRU-first, requests still run and stay correct."""
    models, eng = _prefix_engine(num_pages=12, page_size=4)                 # 11 usable pages
    g = SamplingParams(temperature=0.0)
    prompts = [[i * 10 + j for j in range(9)] for i in range(1, 6)]         # 5 unrelated 9-token prompts: 2 cacheable pages each
    for p in prompts:
        assert eng.generate(p, g, max_tokens=4) == _greedy_oracle(models, p, 4)
    assert eng.table.prefix.evictions > 0
    p = prompts[-1]                                                         # most recent prefix is still resident
    before = eng.stats["prefix_cached_tokens"]
    assert eng.generate(p, g, max_tokens=4) == _greedy_oracle(models, p, 4)
    assert eng.stats["prefix_cached_tokens"] == before + 8


def test_mixed_batches_keep_running_streams_decoding_during_prefill():
    """mixed_batches=True: a step that prefills a new prompt also decodes one token of every running sequence; results are
    identical to the separate-phase schedule and running sequences never stall behind a long prompt."""
    long_prompt = list(range(20, 60))
    outs = {}
    for mixed in (False, True):
        models, eng = _engine(mixed_batches=mixed, max_prefill_tokens=8)        # the long prompt needs 5 prefill chunks
        g = SamplingParams(temperature=0.0)
        first = eng.submit([3, 9, 27, 81], g, max_tokens=12)
        for _ in range(3):
            eng.step()
        produced_before = len(first.output)
        late = eng.submit(long_prompt, g, max_tokens=4)
        stalls = 0
        while late.prefilled < len(late.prompt):
            n0 = len(first.output)
            eng.step()
            stalls += int(len(first.output) == n0 and not first.finished)
        eng.drain()
        outs[mixed] = (first.output, late.output, stalls)
        assert first.output == _greedy_oracle(models, [3, 9, 27, 81], 12) and late.output == _greedy_oracle(models, long_prompt, 4)
        assert produced_before > 0
    assert outs[False][:2] == outs[True][:2]
    assert outs[True][2] <= 1 < outs[False][2]          # separate phases stall the running stream for the whole prefill


def test_scheduler_stress_random_traffic_keeps_invariants():
    """Randomised traffic against the scheduler alone (a fake pipeline returns deterministic tokens): arrivals between steps,
    cancellations, EOS / stop sequences, tight KV pool, prefix cache + mixed batches on.  Every request must terminate with a
    legal reason, never exceed max_tokens, and all pages must be back (free or evictable) at the end."""
    import random

    from mlx_sharding_b200.engine.core import StepOutput

    class FakePipe:
        num_stages = 1

        def submit(self, inp):
            # token = f(sequence id, position): deterministic, occasionally the EOS id 1
            toks = [(sid * 7 + int(c)) % 13 + 1 for sid, c in zip(inp.seq_ids, inp.meta.context_lens.tolist())]
            assert inp.meta.num_tokens == inp.tokens.numel() and max(inp.meta.slot_mapping.tolist()) < 40 * 4
            assert len(set(inp.meta.slot_mapping.tolist())) == inp.meta.num_tokens      # no two tokens share a KV slot
            return StepOutput(toks, [0.0] * len(toks))

        def wait(self, h):
            return h

        def reset(self):
            pass

    for seed in range(6):
        rnd = random.Random(seed)
        eng = LLMEngine(FakePipe(), num_pages=40, page_size=4, num_groups=rnd.choice([1, 2]), max_seqs_per_group=rnd.choice([2, 5]),
                        max_prefill_tokens=rnd.choice([3, 8, 64]), prefix_cache=rnd.random() < 0.7, mixed_batches=rnd.random() < 0.5)
        live, done = [], []
        system = [rnd.randrange(20, 30) for _ in range(9)]
        for it in range(400):
            if rnd.random() < 0.25 and len(live) < 12:
                prompt = (system if rnd.random() < 0.5 else []) + [rnd.randrange(20, 60) for _ in range(rnd.randrange(1, 14))]
                r = eng.submit(prompt, SamplingParams(temperature=0.0), max_tokens=rnd.randrange(1, 12),
                               eos_token_id=1 if rnd.random() < 0.5 else None,
                               stop_id_sequences=[[5, 6]] if rnd.random() < 0.2 else None)
                live.append(r)
            if live and rnd.random() < 0.05:
                rnd.choice(live).cancel()
            eng.step()
            for r in list(live):
                if r.finished:
                    live.remove(r)
                    done.append(r)
        for r in live:
            r.cancel()
        eng.drain()
        done += live
        assert done and all(r.finished for r in done)
        for r in done:
            assert r.error is None or isinstance(r.error, MemoryError)
            if r.error is None:
                assert r.finish_reason in ("stop", "length", "cancelled") and len(r.output) <= r.max_tokens
                if r.finish_reason == "length":
                    assert len(r.output) == r.max_tokens
                if r.finish_reason == "stop":
                    assert r.output[-1] == r.eos_token_id or r.output[-2:] == [5, 6]
        assert not eng.table.pages and eng.table.alloc.num_free == 40 - 1
        if eng.table.prefix is not None:
            assert all(v == 0 for v in eng.table.prefix.refs.values())


def test_streaming_detokenizer_keeps_leading_space_after_newline():
    """SentencePiece-style decoders drop the leading space of the first token they are given; the streaming detokenizer restarts its
    window after every newline and must not lose indentation there."""
    from mlx_sharding_b200.engine.tokenizer import StreamingDetokenizer

    class SpmLike:
        vocab = {1: "def", 2: "▁f():", 3: "\n", 4: "▁▁▁▁return", 5: "▁1", 6: "\n", 7: "▁x"}

        def decode(self, ids):
            s = "".join(self.vocab[i] for i in ids).replace("▁", " ")
            return s[1:] if s.startswith(" ") else s        # Strip(start=1)

    d = StreamingDetokenizer(SpmLike())
    out = ""
    for t in [1, 2, 3, 4, 5, 6, 7]:
        d.add_token(t)
        out += d.last_segment
    d.finalize()
    out += d.last_segment
    assert out == "def f():\n    return 1\n x", repr(out)
    assert d.text == out


def test_max_tokens_zero_generates_nothing():
    from helpers import TINY_LLAMA
    from mlx_sharding_b200.config import ModelConfig
    from mlx_sharding_b200.models import build_stage
    from mlx_sharding_b200.parallel.pipeline import LocalPipeline
    from mlx_sharding_b200.utils.checkpoint import random_state_dict

    cfg = ModelConfig.from_dict(TINY_LLAMA)
    m = build_stage(cfg, cfg.shard(), torch.float32).load_state(dict(random_state_dict(cfg, dtype=torch.float32)))
    eng = LLMEngine(LocalPipeline.from_models([m], 16, 16), 16, 16)
    r = eng.submit([1, 2, 3], SamplingParams(), max_tokens=0)
    assert r.finished and r.finish_reason == "length" and r.output == [] and list(r) == []
    assert eng.table.alloc.num_free == 15 and not eng.has_work()
