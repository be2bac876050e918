"""one fresh-process execution of tests/test_backward_parity_gpu.py::test_gradients_with_attention_in_the_loss (strict),
the first GPU work of the process; exit code 1 + the full mismatch record on a miss (scripts/hunt_flake2.sh loops it)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import test_backward_parity_gpu as t
try:
    t.test_gradients_with_attention_in_the_loss()
except AssertionError as e:
    print("MISMATCH", e, flush=True)
    sys.exit(1)
print("ok")
