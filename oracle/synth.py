"""Synthetic panoramas for the tests: re-export of ``tools/synth_inputs.py`` (the generators hold no projection
arithmetic; the benchmarks import them from ``tools`` so that nothing outside tests / smoke / CPU-baseline legs touches
``oracle/``)."""
import sys
from pathlib import Path

_ROOT = Path(__file__).resolve().parent.parent
if str(_ROOT) not in sys.path:
    sys.path.insert(0, str(_ROOT))
from tools.synth_inputs import coords, make, noise, smooth  # noqa: E402,F401
