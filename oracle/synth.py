"""Re-export of icl_b200.utils.synth (deterministic synthetic parameters / inputs) for the oracle-side tools and tests.
The generator itself lives in the package so that bench.py's GPU arm imports nothing from oracle/."""
from icl_b200.utils.synth import *  # noqa: F401,F403
from icl_b200.utils.synth import _rng, _uniform  # noqa: F401
