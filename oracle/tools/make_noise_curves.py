"""BUILD TOOL - pack the public detector noise curves the benchmark configuration needs
(aLIGO_O4_high_asd.txt for H1/L1, AdV_psd.txt for V1; plus the
design curve aLIGO_ZERO_DET_high_P_psd.txt) from the reference's data directory
(/root/reference/bilby/gw/detector/noise_curves/, data files - not source) into one
binary ``bilby_b200/data/noise_curves.npz`` so they travel to the GPU box.

Run (build container only):  python oracle/tools/make_noise_curves.py
"""
import os
import numpy as np

SRC = "/root/reference/bilby/gw/detector/noise_curves"
CURVES = ["aLIGO_O4_high_asd.txt", "AdV_psd.txt", "aLIGO_ZERO_DET_high_P_psd.txt",
          "aLIGO_ZERO_DET_high_P_asd.txt", "AdV_asd.txt"]


def main():
    out = {}
    for name in CURVES:
        arr = np.genfromtxt(os.path.join(SRC, name))
        out[name + ":frequency"] = arr[:, 0]
        out[name + ":value"] = arr[:, 1]
    here = os.path.dirname(os.path.abspath(__file__))
    dst = os.path.join(here, "..", "..", "bilby_b200", "data", "noise_curves.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
