#!/bin/bash
# Collect the round's GPU evidence on the box (run through gpurun from the repo root):
#   bash tools/collect_profiles.sh TAG      -> gpurun_out/TAG_*  (text only: .ncu-rep files are summarised here and deleted,
#                                              gpurun only copies back 64 MiB)
TAG=${1:-final}
O=gpurun_out
mkdir -p $O
# 1. bench line + per-shape kernel table (CUDA events, warm, inside a step)
timeout 300 python bench.py --steps 8 --warmup 3 --kernel-table $O/${TAG}_kernel_table.md > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
# 2. launch list of ONE step (cold-cache, serialised): per-kernel time + DRAM bytes
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --profile-from-start off --csv --log-file /tmp/${TAG}_launches.csv python bench.py --profile --steps 1 --warmup 3 > /dev/null 2>&1
python tools/summarize_launches.py /tmp/${TAG}_launches.csv $O/${TAG}_launches_step.md > /dev/null 2>&1
gzip -c /tmp/${TAG}_launches.csv > $O/${TAG}_launches_step.csv.gz
# 3. full captures of the dominant kernels, summarised on the box
cap() {  # name, kernel regex, skip, python script, [env]
  local name=$1 rx=$2 skip=$3 script=$4
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o /tmp/ncu_$name -f python $script > /dev/null 2>&1
  (python tools/ncu_summary.py /tmp/ncu_$name.ncu-rep; echo; echo "#### top stall sites (SASS, warp-state samples)"; echo; echo '```';
   python tools/ncu_hotspots.py /tmp/ncu_$name.ncu-rep | head -40; echo '```') > $O/${TAG}_ncu_$name.md 2>&1
  rm -f /tmp/ncu_$name.ncu-rep
}
X2K_CASE="fc1 fwd bias+gelu+gelu'" cap gemm_fc1_gelu gemm_tcgen05 6 tools/profile_gemm.py
X2K_CASE="plain fwd bf16" cap gemm_plain_bf16 gemm_tcgen05 6 tools/profile_gemm.py
X2K_CASE="o-proj bias+drop+res+f32" cap gemm_oproj_drop_res gemm_tcgen05 6 tools/profile_gemm.py
X2K_ATTN_CASE=beit577 cap attn_long_fwd_577 attn_long_fwd 2 tools/profile_attn.py
X2K_ATTN_CASE=beit577 cap attn_long_bwd_577 attn_long_bwd 2 tools/profile_attn.py
X2K_ATTN_CASE=beit cap attn_bwd_beit attn_bwd_kernel 2 tools/profile_attn.py
X2K_ATTN_CASE=cross cap attn_bwd_cross attn_pack_bwd 2 tools/profile_attn.py
X2K_ATTN_CASE=cross cap attn_fwd_cross attn_pack_fwd 2 tools/profile_attn.py
cap layernorm_bwd_dropcast "layernorm_bwd_kernel.*6.*1" 4 "bench.py --profile --steps 1 --warmup 3"
python tools/profile_attn.py > $O/${TAG}_attn_shapes.txt 2>&1
python tools/epilogue_ab.py > $O/${TAG}_gemm_epilogue_ab.txt 2>&1
ls -la $O | grep ${TAG}_
