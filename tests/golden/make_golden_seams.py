"""Golden vectors for the small numeric seams of utils/math.py, from the UNMODIFIED reference (build container).

Run:  python tests/golden/make_golden_seams.py  ->  tests/golden/seams.npz
conv_log (:212-238), fft_convolve_two_children (:241-263, G = 1024 and 1088), non_log_conv (:266-282).
"""

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(REPO, "oracle", "refshim"), "/root/reference"]

import numpy as np  # noqa: E402

from phyclone.utils.math import conv_log, fft_convolve_two_children, non_log_conv  # noqa: E402


def main():
    rng = np.random.default_rng(777)
    out = {}
    for n in (1, 2, 30, 101):
        x, y = np.log(rng.uniform(0.1, 1000, n)), np.log(rng.uniform(0.1, 1000, n))
        out["convlog_x_%d" % n], out["convlog_y_%d" % n] = x, y
        out["convlog_out_%d" % n] = conv_log(x, y, np.zeros(n))
        out["nonlog_out_%d" % n] = non_log_conv(x, y)
    # wide dynamic range: the per-output shift of conv_log keeps terms the per-row shift of non_log_conv floors
    x = -np.arange(60.0) ** 2
    y = -(np.arange(60.0) - 40) ** 2 * 3
    out["convlog_x_wide"], out["convlog_y_wide"] = x, y
    out["convlog_out_wide"] = conv_log(x, y, np.zeros(60))
    out["nonlog_out_wide"] = non_log_conv(x, y)
    for G in (1024, 1088):
        a, b = np.log(rng.uniform(0.1, 1000, (2, G))), np.log(rng.uniform(0.1, 1000, (2, G)))
        out["fft_a_%d" % G], out["fft_b_%d" % G] = a, b
        out["fft_out_%d" % G] = fft_convolve_two_children(a, b)
    np.savez_compressed(os.path.join(HERE, "seams.npz"), **out)
    print("wrote seams.npz", os.path.getsize(os.path.join(HERE, "seams.npz")))


if __name__ == "__main__":
    main()
