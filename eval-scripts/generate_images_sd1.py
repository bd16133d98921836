"""SD1.x slider scale sweep over a prompts CSV — same command line as the reference's eval-scripts/generate_images_sd1.py
(`--scheduler lms` is that script's LMSDiscreteScheduler, :51).  See sliders_b200/eval_sweep.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sliders_b200 import eval_sweep  # noqa: E402

if __name__ == "__main__":
    eval_sweep.main(xl=False)
