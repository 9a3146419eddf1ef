"""Issue-rate microbenchmarks of the instructions the filter kernels are built from (pa_microbench_dpx modes 0-7, run on the
B200 box): the ceilings the roofline of bench.py is quoted against, and the co-issue experiments of DESIGN.md section 5c
(profiles/r2_microbench_pipes.txt)."""
import sys; sys.path.insert(0,'.')
from phyloaln_b200 import capi
c=capi.Context(0)
names=["viaddmnmx_s16x2_relu","vimnmx3_s16x2","ssv_mix_dpx","viaddmnmx_s32","hfma2_relu","ssv_mix_hfma2","hfma2_warps_next_to_dpx_warps(cell ops)","hfma2+dpx_in_every_warp(cell ops)"]
for i,n in enumerate(names):
    g=c.microbench_dpx(i)
    print(f"mode {i} {n}: {g:.0f} Ginstr/s = {g*1e9/148/1.965e9:.1f} thread-instr/clk/SM")
