"""Property tests (hypothesis) of the host-side invariants the kernels rely on: MLX affine packing, the kernel nibble
re-pack, ragged step metadata (slot mapping / packing), the stage partitioner and the paged-KV allocator."""
import torch
from hypothesis import given, settings, strategies as st

from mlx_sharding_b200.config import ModelConfig, ShardSpec
from mlx_sharding_b200.engine.kv_cache import PageAllocator, SequenceTable
from mlx_sharding_b200.ops.meta import BatchMeta
from mlx_sharding_b200.parallel.partition import balanced_split, stage_cost
from mlx_sharding_b200.utils import quant

from helpers import TINY_DSV2, TINY_LLAMA

FAST = settings(max_examples=40, deadline=None)


@FAST
@given(bits=st.sampled_from([2, 4, 8]), rows=st.integers(1, 5), words=st.integers(1, 6), seed=st.integers(0, 2 ** 16))
def test_pack_unpack_roundtrip(bits, rows, words, seed):
    g = torch.Generator().manual_seed(seed)
    codes = torch.randint(0, 1 << bits, (rows, words * (32 // bits)), generator=g, dtype=torch.uint8)
    packed = quant.pack_codes(codes, bits)
    assert packed.dtype == torch.int32 and packed.shape == (rows, words)
    assert torch.equal(quant.unpack_codes(packed, bits), codes)
    # LSB first: code 0 of every word sits in the lowest bits
    assert torch.equal((packed.long() & ((1 << bits) - 1)).to(torch.uint8), codes[:, :: 32 // bits])


@FAST
@given(rows=st.integers(1, 4), words=st.integers(1, 8), seed=st.integers(0, 2 ** 16))
def test_int4_pair_repack_is_a_permutation_of_nibbles(rows, words, seed):
    """Kernel layout: nibble j <- code 2j, nibble 4 + j <- code 2j + 1, so (w >> 4i) & 0x000F000F == (code 2i, code 2i+1)."""
    g = torch.Generator().manual_seed(seed)
    codes = torch.randint(0, 16, (rows, words * 8), generator=g, dtype=torch.uint8)
    rp = quant.repack_int4_pairs(quant.pack_codes(codes, 4)).long() & 0xFFFFFFFF
    c = codes.view(rows, words, 8).long()
    for i in range(4):
        pair = (rp >> (4 * i)) & 0x000F000F
        assert torch.equal(pair & 0xF, c[..., 2 * i]) and torch.equal(pair >> 16, c[..., 2 * i + 1])


@FAST
@given(group=st.sampled_from([32, 64]), bits=st.sampled_from([4, 8]), n_groups=st.integers(1, 3), seed=st.integers(0, 2 ** 16))
def test_affine_quantisation_error_is_bounded_by_the_step(group, bits, n_groups, seed):
    """mx.quantize snaps the scale so that the larger-magnitude edge of the group is exactly representable; the grid may then
    stop up to ~1 step short of the other edge, so the worst case is 1.5 steps there and half a step in the interior."""
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(3, group * n_groups, generator=g)
    wq, s, b = quant.quantize(w, group, bits, out_dtype=torch.float32)
    deq = quant.dequantize(wq, s, b, group, bits)
    step = s.abs().repeat_interleave(group, dim=-1)
    err = (deq - w).abs()
    assert (err <= 1.5 * step + 1e-6).all()
    grp = w.view(3, n_groups, group)
    edge = torch.where(grp.amin(-1).abs() > grp.amax(-1).abs(), grp.amin(-1), grp.amax(-1))
    hit = (deq.view(3, n_groups, group) - edge.unsqueeze(-1)).abs().amin(-1)
    assert (hit <= 1e-6).all()                                            # the dominant edge is reproduced exactly


@st.composite
def ragged_batches(draw):
    page = draw(st.sampled_from([4, 16]))
    n = draw(st.integers(1, 5))
    q = [draw(st.integers(1, 9)) for _ in range(n)]
    ctx = [draw(st.integers(0, 20)) for _ in range(n)]
    pages, nxt = [], 1
    for qi, ci in zip(q, ctx):
        need = (qi + ci + page - 1) // page
        pages.append(list(range(nxt, nxt + need)))
        nxt += need
    return page, q, ctx, pages


@FAST
@given(ragged_batches(), st.sampled_from([0, 8]))
def test_batch_meta_slots_and_packing(batch, pad):
    page, q, ctx, pages = batch
    m = BatchMeta.build(q, ctx, pages, page, pad_blocks_to=pad)
    assert m.num_tokens == sum(q) and m.num_seqs == len(q) and m.max_q_len == max(q)
    t = 0
    seen = set()
    for b, (qi, ci) in enumerate(zip(q, ctx)):
        assert int(m.context_lens[b]) == qi + ci and int(m.last_idx[b]) == t + qi - 1
        for j in range(qi):
            pos = ci + j
            assert int(m.positions[t]) == pos
            slot = pages[b][pos // page] * page + pos % page
            assert int(m.slot_mapping[t]) == slot and slot not in seen       # every token owns a distinct KV slot
            seen.add(slot)
            t += 1
    if pad:
        assert m.block_tables.shape[1] % pad == 0
    m2 = BatchMeta.unpack(m.pack())
    for f in ("positions", "slot_mapping", "cu_seqlens", "context_lens", "last_idx", "block_tables"):
        assert torch.equal(getattr(m, f).int(), getattr(m2, f).int()), f
    assert m.pack().numel() == BatchMeta.packed_size(m.num_tokens, m.num_seqs, m.block_tables.shape[1])


@FAST
@given(cfgd=st.sampled_from([TINY_LLAMA, TINY_DSV2]), layers=st.integers(2, 12), stages=st.integers(1, 9), half=st.booleans())
def test_partition_is_contiguous_complete_and_not_worse_than_even(cfgd, layers, stages, half):
    cfg = ModelConfig.from_dict(dict(cfgd, num_hidden_layers=layers))
    plan = balanced_split(cfg, stages, half_layers=half)
    assert 1 <= len(plan) <= stages
    units = [(i, blk) for sp in plan for i in sp.layers() for blk in "am" if (sp.runs_attn(i) if blk == "a" else sp.runs_mlp(i))]
    assert units == [(i, blk) for i in range(layers) for blk in "am"]
    assert plan[0].is_first and plan[-1].is_last and sum(sp.is_first for sp in plan) == 1 and sum(sp.is_last for sp in plan) == 1
    if len(plan) == stages and stages <= layers:
        even = ShardSpec.even_split(layers, stages)
        assert max(stage_cost(cfg, sp) for sp in plan) <= max(stage_cost(cfg, sp) for sp in even) + 1e-12


@FAST
@given(st.lists(st.tuples(st.integers(0, 3), st.integers(1, 40)), min_size=1, max_size=30), st.sampled_from([4, 16]))
def test_page_allocator_never_double_books(ops, page):
    """Random reserve / release traffic: pages are never shared between live sequences and all come back at the end."""
    alloc = PageAllocator(64)
    free0 = alloc.num_free
    table = SequenceTable(alloc, page)
    live = {}
    for sid, (kind, n) in enumerate(ops):
        if kind == 0 and live:
            victim = sorted(live)[n % len(live)]
            table.release(victim)
            live.pop(victim)
            continue
        need = (n + page - 1) // page
        if need > alloc.num_free:
            continue
        table.add(sid)
        table.reserve(sid, n)
        table.advance(sid, n)
        live[sid] = list(table.pages[sid])
        owned = [p for ps in live.values() for p in ps]
        assert len(owned) == len(set(owned)) and 0 not in owned          # page 0 is the reserved null page
    for sid in list(live):
        table.release(sid)
    assert alloc.num_free == free0


def test_batchmeta_fresh_flag_and_its_host_side_recomputation():
    """``BatchMeta.fresh`` (no sequence has cached context) is set by ``build`` and recomputed by every pipeline stage from the packed
    step block (context length == query length for all sequences) — the hint the DeepSeek prefill fast path keys on."""
    import numpy as np

    from mlx_sharding_b200.engine.sampler import SamplingParams
    from mlx_sharding_b200.ops.meta import BatchMeta
    from mlx_sharding_b200.parallel.graph_decode import HEADER_WORDS, pack_step, unpack_header

    bts = [[1, 2], [3, 4], [5, 6]]
    cases = [([5, 7, 3], [0, 0, 0], True), ([5, 7, 3], [0, 4, 0], False), ([1, 1, 1], [9, 2, 5], False), ([4], [0], True)]
    for q, c0, want in cases:
        m = BatchMeta.build(q, c0, bts[:len(q)], 16)
        assert m.fresh is want and m.to("cpu").fresh is want
        toks = torch.arange(sum(q), dtype=torch.int64)
        wire, lay = pack_step(1, m, toks, [SamplingParams(temperature=0.0)] * len(q), [[]] * len(q), None, max(q) > 1)
        lay2, _ = unpack_header(wire)
        blk = wire[HEADER_WORDS:HEADER_WORDS + lay2.size]
        o = lay2.meta + 6 + 2 * lay2.T
        cu, ctx = blk[o:o + lay2.B + 1], blk[o + lay2.B + 1:o + 2 * lay2.B + 1]
        assert bool(np.array_equal(ctx, cu[1:] - cu[:-1])) is want
