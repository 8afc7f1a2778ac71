"""The end-to-end parity machinery lives in oracle/parity.py (test infrastructure shared with bench.py's parity gate)."""
from oracle.parity import *  # noqa: F401,F403
from oracle.parity import BORDERLINE_PX, Recorder, as_rows, compare_keypoints, compare_stage1, oracle_merge_of_rows, probe, run_sliced_case, to_xywh  # noqa: F401
