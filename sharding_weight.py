#!/usr/bin/env python
"""Offline weight pre-splitter: write one pipeline stage of a checkpoint into its own directory.

Same CLI as the reference script (``sharding_weight.py:74-87``, underscored flags, all required):

    python sharding_weight.py --model <path-or-repo> --output_dir shard_0 \
        --start_layer 0 --end_layer 14 --total_layers 27

Output layout (reference ``sharding_weight.py:26-71``): ``model-{start:05d}-{end:05d}.safetensors``
(+ ``.index.json`` when the source has an index), ``config.json`` with ``start_layer``/``end_layer``
baked in, and every non-weight file (tokenizer etc.) copied over.
"""
import argparse
import os

from mlx_sharding_b200.utils.checkpoint import copy_other_files, save_sharded_weights


def main(argv=None):
    parser = argparse.ArgumentParser(description="Save sharded weights for one pipeline stage")
    parser.add_argument("--model", type=str, required=True, help="Path to the model or HuggingFace repo")
    parser.add_argument("--output_dir", type=str, required=True, help="Directory to save the sharded weights")
    parser.add_argument("--start_layer", type=int, required=True, help="Start layer index (inclusive)")
    parser.add_argument("--end_layer", type=int, required=True, help="End layer index (exclusive)")
    parser.add_argument("--total_layers", type=int, required=True, help="Total number of layers in the model")
    args = parser.parse_args(argv)
    os.makedirs(args.output_dir, exist_ok=True)
    out = save_sharded_weights(args.model, args.output_dir, args.start_layer, args.end_layer, args.total_layers)
    print(f"Sharded weights saved to {out}")
    copy_other_files(args.model, args.output_dir)
    print(f"Other files copied to {args.output_dir}")


if __name__ == "__main__":
    main()
