"""TEST INFRASTRUCTURE ONLY.  The deterministic synthetic weights / videos live in the neutral top-level module
``synth_data`` (bench.py's product arm uses them too and must not import ``oracle/``); this shim keeps the
``oracle.weights`` name the tests and the golden generator use."""
from synth_data import _seed_for, fill_state_dict_, is_generator_key, synth_tensor, synth_video  # noqa: F401
