"""Reference-compatible helpers (``shard/utils.py``): ``load_model`` and the generation-step factory."""
from mlx_sharding_b200.utils.loader import load_model  # noqa: F401
from mlx_sharding_b200.engine.compat import create_generate_step_with_grpc  # noqa: F401
from mlx_sharding_b200.parallel.grpc_compat import message_to_tensor as bytes_message_to_tensor  # noqa: F401
from mlx_sharding_b200.parallel.grpc_compat import tensor_to_message  # noqa: F401
