#!/bin/bash
# Round evidence in one GPU call (single GPU): bench lines of both arms, the ncu launch list of the SAME bench command,
# and `ncu --set full` captures of the dominant kernel and of the decoder.  Outputs land in gpurun_out/ (scratch);
# tools/summarise_ncu.py turns the .ncu-rep files into the text summaries committed under profiles/.
# usage: tools/collect_profiles.sh <tag>        e.g.  gpurun -- 'tools/collect_profiles.sh r02'
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
# the dominant kernel first: bench.py reports its DRAM traffic from the summary of THIS build
ncu --set full --import-source on --clock-control none -k regex:k_bucket_accum -c 1 -f -o $out/${tag}_accum \
    python tools/profile_msm.py 20 1 > $out/${tag}_ncu_accum.log 2>&1
python tools/summarise_ncu.py $out/${tag}_accum.ncu-rep k_bucket_accum > profiles/r02_ncu_full_k_bucket_accum.txt
cp profiles/r02_ncu_full_k_bucket_accum.txt $out/${tag}_ncu_full_k_bucket_accum.txt
python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err
python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_reference_cpu.json 2> $out/${tag}_bench_reference_cpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches_bench_steps2_warmup3.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $out/${tag}_ncu_bench.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_decompress -c 1 -f -o $out/${tag}_decompress \
    python tools/profile_msm.py 20 1 > $out/${tag}_ncu_decompress.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_digit_scatter -c 1 -f -o $out/${tag}_scatter \
    python tools/profile_msm.py 20 1 > $out/${tag}_ncu_scatter.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_tree_level_quad -c 1 -f -o $out/${tag}_leaf \
    python tools/profile_msm.py 20 1 > $out/${tag}_ncu_leaf.log 2>&1
ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum --csv --log-file $out/${tag}_launches_msm_2e20_warm.csv \
    python tools/profile_msm.py 20 4 > /dev/null 2>&1
for l in 10 16; do
  ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum --csv --log-file $out/${tag}_launches_msm_2e${l}_warm.csv \
      python tools/profile_msm.py $l 6 > /dev/null 2>&1
done
tail -c 400 $out/${tag}_bench_1gpu.err
ls -la $out | grep ${tag}_
