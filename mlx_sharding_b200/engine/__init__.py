"""Generation engine: paged KV cache, sampler, scheduler, generation driver, tokenizer utilities."""
