#!/bin/bash
# One `ncu --set full` capture per kernel class on the final library (GPU box, one GPU), after the bench command has
# run once without ncu.  The launches are those of the 4th encoder step of `bench.py --steps 1 --warmup 3`.
set -x
cd $GRAFT_REPO_ROOT
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-incremental"
timeout 300 $B > gpurun_out/ncu_pre.json 2> gpurun_out/ncu_pre.err || exit 1
# GEMM: out_proj(+res), fc1(+GELU), fc2(+res), QKV of layer 12 of the 4th step (103 launches per step: 7 conv/proj, then 4 per layer)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s $((3*103 + 7 + 4*12)) -c 4 -o gpurun_out/r02_gemm -f $B > gpurun_out/ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s $((3*24 + 12)) -c 1 -o gpurun_out/r02_attn -f $B > gpurun_out/ncu_attn.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"conv0_kernel|ln_gelu_bf16_kernel" -s $((3*7)) -c 3 -o gpurun_out/r02_rows -f $B > gpurun_out/ncu_rows.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:layernorm_f32_kernel -s $((3*48 + 24)) -c 1 -o gpurun_out/r02_ln -f $B > gpurun_out/ncu_ln.log 2>&1
for n in gemm attn rows ln; do python tools/ncu_summary.py gpurun_out/r02_$n.ncu-rep > gpurun_out/r02_ncu_${n}_summary.csv 2>/dev/null; done
ls -la gpurun_out/*.ncu-rep
