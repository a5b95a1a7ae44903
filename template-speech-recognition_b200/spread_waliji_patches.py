"""Drop-in for ``spread_waliji_patches`` (template_speech_rec/spread_waliji_patches.pyx:15-59),
which computes exactly what code_parts.spread_patches does."""
from .code_parts import spread_patches as _spread_patches


def spread_waliji_patches(X, spread_0_dim, spread_1_dim, num_parts):
    """X int64 (D0,D1) best-part codes (0 = background) -> uint8 (D0,D1,num_parts) spread map."""
    return _spread_patches(X, spread_0_dim, spread_1_dim, num_parts)
