#!/bin/bash
# One-GPU evidence run (everything the round's profiles/ files come from).  Run on the GPU box from the repo root:
#     MCI_GIT_COMMIT=<short sha> bash tools/evidence_run.sh
# Order matters: the ncu captures come first and their summaries are written to profiles/ (stamped with the hash of the
# kernel sources), so that the bench line produced afterwards can carry roofline.traffic / limiters_ncu for exactly the
# kernels it times.  Everything is copied to gpurun_out/ (the only directory that travels back).
set -u
mkdir -p gpurun_out/profiles
O=gpurun_out
# (0) the GPU test suite on the same build
python -X faulthandler -m pytest tests -m gpu -q --tb=short -rf > $O/r2_pytest_gpu.log 2>&1
tail -3 $O/r2_pytest_gpu.log | cut -c1-200
# (1) ncu --set full of the four fill kernels.  sweep_fill launches (fixed, twin) pairs per iteration; iterations >= 3
#     run the fixed-point kernel on C2, so skip the first launches
ncu --set full --clock-control none --import-source on -k regex:vegas_pool_kernel -s 6 -c 4 -f -o $O/r2_prof_fill python tools/sweep_fill.py gauss10 vegas 26 8192 0 6 > $O/r2_prof_fill.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vegasmc_kernel -s 4 -c 1 -f -o $O/r2_prof_bubble python tools/sweep_fill.py bubble vegasmc 25 9216 0 6 > $O/r2_prof_bubble.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mcmc_kernel -s 4 -c 1 -f -o $O/r2_prof_mcmc python tools/sweep_fill.py product4 mcmc 22 9216 0 6 > $O/r2_prof_mcmc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vegas_generic_kernel -s 8 -c 2 -f -o $O/r2_prof_comp python tools/sweep_fill.py genz16_comp vegas 22 8192 0 6 > $O/r2_prof_comp.log 2>&1
# (2) their summaries: units per launch = 16 blocks x 2^k samples (the sweep_fill sizes above)
python tools/make_ncu_json.py $O/r2_prof_fill.ncu-rep vegas_pool_kernel $((16 << 26)) profiles/r2_fill_ncu.json > /dev/null 2> $O/r2_make_ncu.err
python tools/make_ncu_json.py $O/r2_prof_bubble.ncu-rep vegasmc_kernel $((16 << 25)) profiles/r2_bubble_ncu.json > /dev/null 2>> $O/r2_make_ncu.err
python tools/make_ncu_json.py $O/r2_prof_mcmc.ncu-rep mcmc_kernel 73819750 profiles/r2_mcmc_c5_ncu.json > /dev/null 2>> $O/r2_make_ncu.err
python tools/make_ncu_json.py $O/r2_prof_comp.ncu-rep vegas_generic_kernel $((16 << 22)) profiles/r2_composite_ncu.json > /dev/null 2>> $O/r2_make_ncu.err
python tools/ncu_summary.py $O/r2_prof_fill.ncu-rep > profiles/r2_ncu_fill_fixed.txt 2>> $O/r2_make_ncu.err
python tools/ncu_summary.py $O/r2_prof_bubble.ncu-rep > profiles/r2_ncu_vegasmc_bubble.txt 2>> $O/r2_make_ncu.err
python tools/ncu_summary.py $O/r2_prof_mcmc.ncu-rep > profiles/r2_ncu_mcmc_product4.txt 2>> $O/r2_make_ncu.err
python tools/ncu_summary.py $O/r2_prof_comp.ncu-rep > profiles/r2_ncu_generic_composite16.txt 2>> $O/r2_make_ncu.err
# (3) the bench line (both arms) and the launch list of the same command
python bench.py > profiles/r2_bench_n1.json 2> $O/r2_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > profiles/r2_bench_reference_arm.json 2>> $O/r2_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file profiles/r2_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/r2_bench_under_ncu.log 2>&1
# (4) BASELINE configs end to end, the reference's own timing workload, small-job latency, the kernel table
python tools/run_configs.py > profiles/r2_baseline_configs.jsonl 2> $O/r2_baseline_configs.err
python tools/cuba_benchmark.py > profiles/r2_cuba.jsonl 2> $O/r2_cuba.err
python tests/perf_latency_sweep.py > profiles/r2_time_to_solution.jsonl 2> $O/r2_latency.err
python tools/vegas_benchmarks.py > profiles/r2_vegas_benchmarks.jsonl 2> $O/r2_vegas_benchmarks.err
F=profiles/r2_fill_kernels_final.jsonl
rm -f $F
for c in "gauss10 vegas 26" "genz16_pool vegas 24" "product4 vegas 24"; do python tools/sweep_fill.py $c 8192 0,4,5,n 8 >> $F 2>> $O/r2_fill_final.err; done
for c in "genz16_comp vegas 22" "genz6_comp vegas 22" "lattice_comp vegas 22" "bubble vegas 22"; do python tools/sweep_fill.py $c 8192 0,2,5 8 >> $F 2>> $O/r2_fill_final.err; done
for c in "product4 mcmc 22" "bubble vegasmc 22" "bubble mcmc 22" "bubble_fermik mcmc 22" "sphere2 mcmc 22" "sphere2 vegasmc 22" "product4 vegasmc 22"; do python tools/sweep_fill.py $c 9216 0 8 >> $F 2>> $O/r2_fill_final.err; done
# (5) compute-sanitizer over one small problem per kernel family
(compute-sanitizer --tool memcheck python tools/sanitize_smoke.py; compute-sanitizer --tool racecheck python tools/sanitize_smoke.py) > profiles/r2_compute_sanitizer.txt 2>&1
cp profiles/r2_*ncu.json profiles/r2_ncu_*.txt profiles/r2_bench_n1.json profiles/r2_bench_reference_arm.json profiles/r2_launches_bench_steps2.csv \
   profiles/r2_baseline_configs.jsonl profiles/r2_cuba.jsonl profiles/r2_time_to_solution.jsonl profiles/r2_vegas_benchmarks.jsonl $F profiles/r2_compute_sanitizer.txt $O/profiles/
tail -c 300 profiles/r2_bench_n1.json; echo; tail -3 profiles/r2_compute_sanitizer.txt; cat $O/r2_make_ncu.err | tail -5
