#!/usr/bin/env python
"""Writes the SASS evidence under profiles/sass/ from the built library (no GPU needed):
  <kernel>.sass.txt   full listing of the headline kernels (cuobjdump -sass -fun ...)
  <kernel>.hist.txt   opcode histogram + the instructions that prove the memory path (UBLKCP = cp.async.bulk TMA
                      bulk copy, SYNCS = mbarrier, LDS/STS, LDG, DFMA/DMUL/DADD, MUFU) for every main kernel
    python profiles/make_sass_listing.py
"""
import collections
import os
import re
import subprocess

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
LIB = os.path.join(ROOT, "bilby_b200", "_lib", "libbilby_b200.so")
OUT = os.path.join(ROOT, "profiles", "sass")
KERNELS = {   # label -> (mangled-name regex, keep full listing)
    "k0_prologue_imrphenomd": (r"_Z18bb_prologue_kernelILi0E", False),
    "k0_prologue_taylorf2": (r"_Z18bb_prologue_kernelILi1E", False),
    "k1_inner_product_3det_imrphenomd": (r"_Z23bb_inner_product_kernelILi3ELi0ELb0E", True),
    "k1_inner_product_3det_imrphenomd_cal": (r"_Z23bb_inner_product_kernelILi3ELi0ELb1E", False),
    "k1_inner_product_3det_taylorf2": (r"_Z23bb_inner_product_kernelILi3ELi1ELb0E", False),
    "k3_epilogue": (r"_Z23bb_epilogue_coef_kernel", False),
    "k4_time_marg_3det_imrphenomd_cal": (r"_Z19bb_time_marg_kernelILi3ELi0ELb1E", True),
    "k4a_series_fill_3det_imrphenomd_cal": (r"_Z21bb_series_fill_kernelILi3ELi0ELb1E", True),
    "k4b_series_fft": (r"_Z20bb_series_fft_kernel", False),
    "k4r_series_fine": (r"_Z21bb_series_fine_kernel", False),
    "kr_recon_time": (r"_Z20bb_recon_time_kernel", False),
    "kr_recon_distance_phase": (r"_Z30bb_recon_distance_phase_kernel", False),
    "k5_relbin_3det_taylorf2": (r"_Z16bb_relbin_kernelILi3ELi1ELb0ELb1E", False),
    "k5_multiband_3det_taylorf2": (r"_Z16bb_relbin_kernelILi3ELi1ELb0ELb0E", False),
    "k5t_relbin_time_marg_3det_imrphenomd": (r"_Z26bb_relbin_time_marg_kernelILi3ELi0ELb0E", False),
    "k6_roq_3det_taylorf2": (r"_Z13bb_roq_kernelILi3ELi1ELb0E", True),
    "kx_exchange_signal": (r"_Z25bb_exchange_signal_kernel", True),
    "kx_exchange_epilogue": (r"_Z27bb_exchange_epilogue_kernel", False),
    "kc_calmarg_select": (r"_Z24bb_calmarg_select_kernel", False),
    "k7_roq_hlinear_3det_taylorf2": (r"_Z21bb_roq_hlinear_kernelILi3ELi1ELb0E", False),
    "k7_roq_time_marg_3det": (r"_Z23bb_roq_time_marg_kernelILi3E", False),
    "kt_distance_table": (r"_Z24bb_distance_table_kernel", False),
    "kg_gemm_complex_dmma": (r"_Z17bb_gemm_nt_kernelILb1E", True),
    "kg_gemm_real_dmma": (r"_Z17bb_gemm_nt_kernelILb0E", False),
}


def main():
    os.makedirs(OUT, exist_ok=True)
    names = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.findall(r"Function : (\S+)", names)
    for label, (pat, full) in KERNELS.items():
        match = [f for f in funcs if re.match(pat, f)]
        if not match:
            print("missing", label)
            continue
        fn = match[0]
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn, LIB], capture_output=True, text=True).stdout
        lines = [ln for ln in sass.splitlines() if re.search(r"/\*[0-9a-f]{4,5}\*/", ln)]
        ops = collections.Counter()
        for ln in lines:
            m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
            if m:
                ops[m.group(1).split(".")[0]] += 1
        with open(os.path.join(OUT, label + ".hist.txt"), "w") as fh:
            fh.write(f"# {fn}\n# {len(lines)} SASS instructions (static count); cuobjdump -sass -fun of {os.path.basename(LIB)}\n")
            fp64 = sum(ops[k] for k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
            fh.write(f"# FP64 pipe instructions (DFMA+DMUL+DADD+DSETP): {fp64} = {100.0 * fp64 / max(1, len(lines)):.1f} % of the listing\n")
            for op, c in ops.most_common():
                fh.write(f"{op:12s} {c}\n")
            proof = [ln.strip() for ln in lines if re.search(r"UBLKCP|SYNCS|UTMA|LDGSTS|HMMA|DMMA", ln)]
            if proof:
                fh.write("\n# bulk-copy / mbarrier / tensor instructions\n")
                for ln in proof[:40]:
                    fh.write(ln + "\n")
        if full:
            with open(os.path.join(OUT, label + ".sass.txt"), "w") as fh:
                # address + instruction only (the 128-bit encodings are dropped to keep the file small)
                for ln in lines:
                    fh.write(re.sub(r"\s*/\* 0x[0-9a-f]{16} \*/\s*$", "", ln).rstrip() + "\n")
        print(label, len(lines), "instructions")


if __name__ == "__main__":
    main()
