"""Reference-compatible helpers (``shard/utils.py``): ``load_model`` and the generation-step factory."""
from mlx_sharding_b200.utils.loader import load_model  # noqa: F401
from mlx_sharding_b200.engine.compat import create_generate_step_with_grpc  # noqa: F401
from mlx_sharding_b200.parallel.grpc_compat import message_to_tensor as bytes_message_to_tensor  # noqa: F401
from mlx_sharding_b200.parallel.grpc_compat import tensor_to_message  # noqa: F401


# The reference's wire helpers (shard/utils.py:71-109 upstream), for scripts that import them by name -------------------------------
import numpy as _np  # noqa: E402
import torch as _torch  # noqa: E402

from mlx_sharding_b200.parallel.grpc_compat import DTYPES as _DTYPES  # noqa: E402
from mlx_sharding_b200.parallel.grpc_compat import message_to_tensor as _m2t  # noqa: E402


def tensor_to_bytes(tensor) -> bytes:
    """Raw host bytes of a tensor (reference utils.py:88-90); bf16 travels as its 16-bit pattern."""
    t = tensor.detach().cpu().contiguous()
    return (t.view(_torch.int16) if t.dtype == _torch.bfloat16 else t).numpy().tobytes()


def bytes_to_tensor(byte_data: bytes, dtype_str: str):
    """Inverse of ``tensor_to_bytes`` for the dtype spellings the reference accepts ("mlx.core.float16", ..., reference :93-109)."""
    key = dtype_str.split(".")[-1]
    if key not in _DTYPES:
        raise ValueError(f"Unsupported dtype: {dtype_str}")
    dt = _DTYPES[key]
    if dt == _torch.bfloat16:
        return _torch.from_numpy(_np.frombuffer(byte_data, dtype=_np.int16).copy()).view(_torch.bfloat16)
    npdt = {_torch.float32: _np.float32, _torch.float16: _np.float16, _torch.int32: _np.int32, _torch.int64: _np.int64}[dt]
    return _torch.from_numpy(_np.frombuffer(byte_data, dtype=npdt).copy())


def send_tensor(stub, tensor):
    """``stub.SendTensor`` with a tensor payload (reference :71-76); ``stub`` is a ``grpc_compat.StageStub`` or a generated stub."""
    if hasattr(stub, "_send"):          # mlx_sharding_b200.parallel.grpc_compat.StageStub
        return stub._send(tensor_to_message(tensor), timeout=stub.timeout)
    return stub.SendTensor(tensor_to_message(tensor))


def response_to_mlx_array(response):
    """Tensor of a ``TensorResponse`` or ``None`` on failure (reference :79-85 swallows errors the same way)."""
    try:
        if not response.success:
            return None
        return _m2t(response.tensor)
    except Exception:  # noqa: BLE001
        return None
