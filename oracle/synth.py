"""oracle/synth.py -- re-export of the synthetic data generator (memotr_b200/synthetic.py: pure tensor generation from
seeds, no model arithmetic) under its historical name, so checker code reads `from oracle import synth`."""
from memotr_b200.synthetic import *  # noqa: F401,F403
from memotr_b200.synthetic import _gen  # noqa: F401
