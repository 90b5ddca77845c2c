"""Import the real reference (``/root/reference``) when it is present.

TEST INFRASTRUCTURE ONLY.  The reference is pure Python on top of torch; the
only import-time dependency missing in this image is ``soundfile``
(``diffsptk/utils/public.py:18``), so an empty stub module is put on the path.
``/root/reference`` exists in the build container only -- never on the GPU box
-- so callers must treat ``load_reference() is None`` as "skip".
"""

from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DIFFSPTK_REFERENCE_ROOT", "/root/reference")


def load_reference():
    """Return the imported ``diffsptk`` package of the reference, or None."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "diffsptk")):
        return None
    if "soundfile" not in sys.modules:
        try:
            import soundfile  # noqa: F401
        except Exception:
            sys.modules["soundfile"] = types.ModuleType("soundfile")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    try:
        import diffsptk  # type: ignore
    except Exception:
        return None
    return diffsptk
