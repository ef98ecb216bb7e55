#!/bin/sh
# final_profiles.sh : the measurement pass whose outputs are summarised under profiles/r02/ (run on the GPU box through gpurun).
# Bench lines first (no profiler attached), then the ncu launch list and the full captures of the dominant kernels.
mkdir -p gpurun_out
O=gpurun_out
python bench.py --steps 10 --warmup 3 > $O/p_bench1.json 2> $O/p_bench1.err
python bench.py --config 3 --steps 3 --warmup 2 > $O/p_bench3.json 2> $O/p_bench3.err
python bench.py --config 4 --steps 5 --warmup 3 > $O/p_bench4.json 2> $O/p_bench4.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/p_ref1.json 2> $O/p_ref1.err
python bench.py --impl reference --steps 1 --warmup 0 --ref-fraction 1 > $O/p_ref1_full.json 2> $O/p_ref1_full.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/p_launches.csv \
    python bench.py --steps 2 --warmup 1 --spectra 128 --skip-cpu > $O/p_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sample_likelihood_kernel -s 4 -c 4 -f -o $O/p_likelihood \
    python bench.py --steps 2 --warmup 1 --spectra 128 --skip-cpu > $O/p_ncu_lk.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:voigt_profile_kernel -s 1 -c 1 -f -o $O/p_voigt3 \
    python bench.py --steps 2 --warmup 1 --spectra 128 --skip-cpu > $O/p_ncu_v3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:voigt_profile_kernel -s 1 -c 1 -f -o $O/p_voigt31 \
    python bench.py --config 3 --steps 1 --warmup 1 --spectra 32 --batch 32 --skip-cpu > $O/p_ncu_v31.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:zqso_likelihood_kernel_v2 -c 1 -f -o $O/p_zqso \
    python bench.py --config 4 --steps 1 --warmup 0 --spectra 128 --skip-cpu > $O/p_ncu_zq.log 2>&1
ls -la $O/p_* | awk '{print $5, $9}'
