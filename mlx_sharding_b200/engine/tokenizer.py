"""Tokenizer utilities: HF tokenizer wrapper with an incremental UTF-8-safe detokenizer, and a tiny
self-contained byte-level tokenizer for the offline GPU box / tests.

Reference surface (used at shard/openai_api.py:109-116,362-363,444-445 and generate.py:91-95):
``load_tokenizer(path, config)`` -> wrapper with ``.detokenizer`` exposing
``reset() / add_token(t) / finalize() / .text / .last_segment`` and pass-through
``encode / decode / apply_chat_template / eos_token_id / chat_template``.
"""
from __future__ import annotations

import json
import os
from typing import List, Optional

_REPLACEMENT = "�"


class StreamingDetokenizer:
    """Incremental detokenizer.

    Tokens are decoded in a sliding window that restarts at every newline, so the cost per token is
    bounded; text is only released once it no longer ends in an incomplete UTF-8 sequence (which
    decodes to U+FFFD), so multi-token characters never reach the client half-finished.
    """

    def __init__(self, tokenizer):
        self._tok = tokenizer
        self.reset()

    def reset(self):
        self.tokens: List[int] = []
        self._window: List[int] = []     # tokens of the segment being decoded
        self._committed = ""             # text of finished segments
        self._window_text = ""           # released text of the current segment
        self.offset = 0                  # how much of .text was already handed out
        # Context for the next segment: the last token of the previous one.  SentencePiece-style decoders strip the leading
        # space of the *first* token they are given (Llama-2 / Mistral ``Strip(start=1)``), so a segment decoded on its own loses
        # indentation after every newline; decoding ``[ctx] + window`` and dropping ``decode([ctx])`` keeps it.
        self._ctx: List[int] = []
        self._ctx_text = ""

    @property
    def text(self) -> str:
        return self._committed + self._window_text

    def _decode(self, ids) -> str:
        try:
            return self._tok.decode(ids)
        except Exception:
            return ""

    def add_token(self, token: int):
        self.tokens.append(int(token))
        self._window.append(int(token))
        t = self._decode_window()
        if t.endswith(_REPLACEMENT):
            return  # incomplete multi-byte char: hold back
        self._window_text = t
        if t.endswith("\n"):
            self._committed += t
            self._ctx = self._window[-1:]
            self._ctx_text = self._decode(self._ctx)
            self._window, self._window_text = [], ""

    def _decode_window(self) -> str:
        if not self._ctx:
            return self._decode(self._window)
        t = self._decode(self._ctx + self._window)
        return t[len(self._ctx_text):] if t.startswith(self._ctx_text) else self._decode(self._window)

    def finalize(self):
        if self._window:
            self._window_text = self._decode_window()
        self._committed += self._window_text
        self._window, self._window_text = [], ""

    @property
    def last_segment(self) -> str:
        """Text produced since the previous ``last_segment`` read."""
        t = self.text
        seg = t[self.offset:]
        self.offset = len(t)
        return seg


class TokenizerWrapper:
    """Pass-through wrapper that adds ``.detokenizer`` (upstream ``TokenizerWrapper``)."""

    def __init__(self, tokenizer, detokenizer_class=StreamingDetokenizer):
        self._tokenizer = tokenizer
        self._detokenizer = detokenizer_class(self)

    @property
    def detokenizer(self) -> StreamingDetokenizer:
        return self._detokenizer

    def new_detokenizer(self) -> StreamingDetokenizer:
        """A private detokenizer (the shared one is not safe with concurrent requests)."""
        return StreamingDetokenizer(self)

    def decode(self, ids, **kw) -> str:
        n = len(self._tokenizer)
        ids = [int(i) for i in ids if 0 <= int(i) < n]  # random-init models can emit ids past the vocab
        return self._tokenizer.decode(ids, **kw)

    def __getattr__(self, name):
        if name in ("_tokenizer", "_detokenizer"):
            raise AttributeError(name)
        return getattr(self._tokenizer, name)

    def __setattr__(self, name, value):
        if name in ("_tokenizer", "_detokenizer"):
            super().__setattr__(name, value)
        else:
            setattr(self._tokenizer, name, value)


def load_tokenizer(model_path: str, tokenizer_config_extra: Optional[dict] = None) -> TokenizerWrapper:
    from transformers import AutoTokenizer

    kw = {k: v for k, v in (tokenizer_config_extra or {}).items() if v is not None}
    return TokenizerWrapper(AutoTokenizer.from_pretrained(str(model_path), **kw))


# --------------------------------------------------------------------------------------------------
# Self-contained byte-level tokenizer (no network, no merges): 256 byte tokens + specials.
# --------------------------------------------------------------------------------------------------
DEFAULT_CHAT_TEMPLATE = (
    "{% for message in messages %}"
    "{{ '<|' + message['role'] + '|>' + message['content'] + '<|end|>' }}"
    "{% endfor %}"
    "{% if add_generation_prompt %}{{ '<|assistant|>' }}{% endif %}"
)


def write_byte_tokenizer(path: str, vocab_size: int = 320, max_filler: int = 4096) -> str:
    """Write ``tokenizer.json`` + ``tokenizer_config.json`` for a byte-level tokenizer.

    ids 0..255 = bytes (GPT-2 byte<->unicode alphabet), then ``<|bos|> <|eos|> <|user|> <|assistant|>
    <|system|> <|end|>``, then filler tokens up to ``min(vocab_size, max_filler)``.
    """
    from tokenizers import Tokenizer, decoders, models, pre_tokenizers
    from tokenizers.pre_tokenizers import ByteLevel

    os.makedirs(path, exist_ok=True)
    alphabet = sorted(ByteLevel.alphabet())
    # GPT-2 byte->unicode order: make id == byte value
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("\xa1"), ord("\xac") + 1)) + \
        list(range(ord("\xae"), ord("\xff") + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    byte_to_char = {b: chr(c) for b, c in zip(bs, cs)}
    assert set(byte_to_char.values()) == set(alphabet)
    vocab = {byte_to_char[b]: b for b in range(256)}
    specials = ["<|bos|>", "<|eos|>", "<|user|>", "<|assistant|>", "<|system|>", "<|end|>"]
    tok = Tokenizer(models.BPE(vocab=vocab, merges=[]))
    tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)
    tok.decoder = decoders.ByteLevel()
    tok.add_special_tokens(specials)
    n_fill = max(0, min(vocab_size, max_filler) - 256 - len(specials))
    if n_fill:
        tok.add_tokens([f"<|extra_{i}|>" for i in range(n_fill)])
    tok.save(os.path.join(path, "tokenizer.json"))
    cfg = {
        "tokenizer_class": "PreTrainedTokenizerFast",
        "bos_token": "<|bos|>", "eos_token": "<|eos|>", "pad_token": "<|eos|>",
        "chat_template": DEFAULT_CHAT_TEMPLATE,
        "clean_up_tokenization_spaces": False,
        "model_max_length": 1 << 20,
    }
    with open(os.path.join(path, "tokenizer_config.json"), "w") as f:
        json.dump(cfg, f, indent=2)
    with open(os.path.join(path, "special_tokens_map.json"), "w") as f:
        json.dump({"bos_token": "<|bos|>", "eos_token": "<|eos|>", "pad_token": "<|eos|>"}, f)
    return path
