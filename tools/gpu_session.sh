#!/bin/bash
# One gpurun call = several measurement stages; every stage logs under gpurun_out/.
#   tools/gpu_session.sh <tag> stage [stage...]      stages: probe tests bench bench_replay launches ncu_full report
set -u
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.csv" 2>&1
for stage in "$@"; do
  echo "=== $stage $(date +%T)"
  case $stage in
    probe) ./tools/probes/cond_graph_probe.bin > "$OUT/probe.log" 2>&1; tail -2 "$OUT/probe.log" ;;
    tests) timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > "$OUT/tests.log" 2>&1; tail -25 "$OUT/tests.log" ;;
    tests_new) timeout 1200 python -m pytest tests/test_gpu_trace.py tests/test_gpu_edges.py -m gpu -q --tb=short -p no:cacheprovider > "$OUT/tests_new.log" 2>&1; tail -25 "$OUT/tests_new.log" ;;
    bench) timeout 900 python bench.py --steps 5 --warmup 3 > "$OUT/bench.json" 2> "$OUT/bench.err"; tail -c 600 "$OUT/bench.json"; tail -3 "$OUT/bench.err" ;;
    bench_quick) timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu ${BENCH_ARGS:-} > "$OUT/bench_quick.json" 2> "$OUT/bench_quick.err"; python tools/bench_brief.py "$OUT/bench_quick.json"; tail -3 "$OUT/bench_quick.err" ;;
    bench_replay) PSI_FIT_LOOP=replay timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu > "$OUT/bench_replay.json" 2> "$OUT/bench_replay.err"; python tools/bench_brief.py "$OUT/bench_replay.json"; tail -3 "$OUT/bench_replay.err" ;;
    bench_u*) U=${stage#bench_u}; PSI_FIT_UNROLL=$U timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu > "$OUT/$stage.json" 2> "$OUT/$stage.err"; python tools/bench_brief.py "$OUT/$stage.json" | head -1; tail -3 "$OUT/$stage.err" ;;
    tests_k) timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "$TESTS_K" > "$OUT/tests_k.log" 2>&1; tail -25 "$OUT/tests_k.log" ;;
    train_s2) timeout 900 python bench.py --workload train_s2 --steps 10 --warmup 3 > "$OUT/train_s2.json" 2> "$OUT/train_s2.err"; tail -c 1200 "$OUT/train_s2.json"; tail -3 "$OUT/train_s2.err" ;;
    train_s2_off) timeout 900 python bench.py --workload train_s2 --geometry off --steps 10 --warmup 3 > "$OUT/train_s2_off.json" 2> "$OUT/train_s2_off.err"; tail -c 700 "$OUT/train_s2_off.json"; tail -3 "$OUT/train_s2_off.err" ;;
    rooms) timeout 900 python bench.py --workload rooms --steps 2 --warmup 1 > "$OUT/rooms.json" 2> "$OUT/rooms.err"; tail -c 1200 "$OUT/rooms.json"; tail -3 "$OUT/rooms.err" ;;
    mg_*) W=${stage#mg_}; NG=${NGPUS:-2}; case $W in fit) EXTRA="--steps 5 --warmup 3";; train_s2) EXTRA="--workload train_s2 --steps 10 --warmup 3";; rooms) EXTRA="--workload rooms --steps 2 --warmup 1";; esac;
       timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG $EXTRA > "$OUT/$stage.n$NG.json" 2> "$OUT/$stage.n$NG.err"; grep '^{' "$OUT/$stage.n$NG.json" | tail -1 | cut -c1-1500; grep -i "nranks\|error\|Traceback" "$OUT/$stage.n$NG.err" | head -5 ;;
    bench_s*) NS=${stage#bench_s}; PSI_FIT_STREAMS=$NS timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu > "$OUT/$stage.json" 2> "$OUT/$stage.err"; python tools/bench_brief.py "$OUT/$stage.json" | head -1; tail -3 "$OUT/$stage.err" ;;
    bench_pdlmax*) M=${stage#bench_pdlmax}; PSI_PDL=3 PSI_PDL_MAX=$M timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu > "$OUT/$stage.json" 2> "$OUT/$stage.err"; python tools/bench_brief.py "$OUT/$stage.json" 2>/dev/null | head -1 ;;
    bench_vo0) PSI_FIT_VORDER=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu > "$OUT/bench_vo0.json" 2> "$OUT/bench_vo0.err"; python tools/bench_brief.py "$OUT/bench_vo0.json" 2>/dev/null | head -12 ;;
    bench_pdl4) PSI_PDL=4 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu > "$OUT/bench_pdl4.json" 2> "$OUT/bench_pdl4.err"; python tools/bench_brief.py "$OUT/bench_pdl4.json" 2>/dev/null | head -1 ;;
    bench_pdl3) PSI_PDL=3 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu > "$OUT/bench_pdl3.json" 2> "$OUT/bench_pdl3.err"; python tools/bench_brief.py "$OUT/bench_pdl3.json" 2>/dev/null | head -1 ;;
    bench_pdl2) PSI_PDL=2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu > "$OUT/bench_pdl2.json" 2> "$OUT/bench_pdl2.err"; python tools/bench_brief.py "$OUT/bench_pdl2.json" 2>/dev/null | head -1 ;;
    bench_pdl) PSI_PDL=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu > "$OUT/bench_pdl.json" 2> "$OUT/bench_pdl.err"; python tools/bench_brief.py "$OUT/bench_pdl.json" | head -1 ;;
    bench_lbfgs) timeout 600 python bench.py --optimizer lbfgs --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu > "$OUT/bench_lbfgs.json" 2> "$OUT/bench_lbfgs.err"; python tools/bench_brief.py "$OUT/bench_lbfgs.json" | head -12; tail -3 "$OUT/bench_lbfgs.err" ;;
    benchlib_*) V=${stage#benchlib_}; PSI_B200_LIB=$PWD/psi-release_b200/lib/variants/libpsi_b200_$V.so timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu ${BENCH_ARGS:-} > "$OUT/$stage.json" 2> "$OUT/$stage.err"; python tools/bench_brief.py "$OUT/$stage.json" | head -12; tail -2 "$OUT/$stage.err" ;;
    testlib_*) V=${stage#testlib_}; PSI_B200_LIB=$PWD/psi-release_b200/lib/variants/libpsi_b200_$V.so timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "${TESTS_K:-golden or baseline_size or fused_iteration}" > "$OUT/$stage.log" 2>&1; tail -4 "$OUT/$stage.log" ;;
    bench_bf3) PSI_LBS_GEMM=bf3 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu > "$OUT/bench_bf3.json" 2> "$OUT/bench_bf3.err"; python tools/bench_brief.py "$OUT/bench_bf3.json" 2>/dev/null | head -12; tail -3 "$OUT/bench_bf3.err" ;;
    bench_ref) timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"; tail -c 600 "$OUT/bench_ref.json" ;;
    launches) PSI_FIT_LOOP=replay timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 90 --csv --log-file "$OUT/launches.csv" python bench.py --steps 1 --warmup 3 --iters 30 --no-cpu-baseline --no-reference-gpu > "$OUT/launches.log" 2>&1; python tools/ncu_summary.py launches "$OUT/launches.csv" | tail -30 ;;
    launches_warm) PSI_FIT_LOOP=replay timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 400 -c 65 --csv --log-file "$OUT/launches_warm.csv" python bench.py --steps 1 --warmup 3 --iters 30 --no-cpu-baseline --no-reference-gpu > "$OUT/launches_warm.log" 2>&1; python tools/ncu_summary.py launches "$OUT/launches_warm.csv" | tail -20 ;;
    ncu_full) timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:${NCU_KERNELS:-nn_index_group|lbs_vertex_bwd|lbs_skin_fwd|lbs_blend_fwd_tc5|lbs_dcoef_tc5}" -s ${NCU_SKIP:-60} -c ${NCU_COUNT:-10} -o "$OUT/prof" python bench.py --steps 1 --warmup 3 --iters 30 --no-cpu-baseline --no-reference-gpu > "$OUT/ncu_full.log" 2>&1; ls -la "$OUT" ;;
    report) timeout 900 python tools/parity_report.py > "$OUT/parity_report.json" 2> "$OUT/parity_report.err"; cat "$OUT/parity_report.json"; tail -3 "$OUT/parity_report.err" ;;
    smoke) timeout 600 python -c 'import __graft_entry__ as g; g.smoke()' > "$OUT/smoke.log" 2>&1; tail -3 "$OUT/smoke.log" ;;
    *) echo "unknown stage $stage" ;;
  esac
done
echo "=== done $(date +%T)"
