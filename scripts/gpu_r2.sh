#!/bin/bash
# Round-2 GPU call driver: scripts/gpu_r2.sh <stage> [...]   (run under gpurun; everything lands in gpurun_out/)
#   tests      pytest -m gpu (all failures shown, no -x)
#   hop        stencil microbenchmark: plain / residual, unit / variable coefficients
#   bench      default bench.py (headline + parity + other workloads)
#   quick      bench.py without cpu baseline / secondary workloads (1 step)
#   launches   ncu launch list of the quick bench
#   ncufull    ncu --set full of the transfer + residual-stencil kernels
#   dist       multi-GPU tests + bench at N = $NGPU (gpurun --gpus N), one bench per entry of $DIST_CFGS
#   knobs      quick bench per entry of $KNOB_CFGS (environment knobs), same box
#   ab         same box, interleaved: the library of an earlier commit (build/prev) against the current one
#   kbt / small / q256 / hopab   standalone transfer-kernel harness, small-solver grid sweep, 256^3 with / without virtual slabs, stencil vs round 1
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
TAG=${TAG:-r02}
for stage in "$@"; do
case $stage in
tests)
    timeout 1500 python -m pytest tests -m gpu -q -s > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
    grep -E "passed|failed|error|parity" $O/${TAG}_pytest_gpu.log | tail -15 ;;
hop)
    : > $O/${TAG}_hop.txt
    for sz in 512x512x512 256x256x256 64x512x512; do
        timeout 200 python scripts/hop_bench.py 2 $sz >> $O/${TAG}_hop.txt 2>&1
        timeout 200 python scripts/hop_bench.py 2 --residual $sz >> $O/${TAG}_hop.txt 2>&1
    done
    HOP_VAR=1 timeout 200 python scripts/hop_bench.py 2 512x256x512 >> $O/${TAG}_hop.txt 2>&1
    HOP_VAR=1 timeout 200 python scripts/hop_bench.py 2 --residual 512x256x512 >> $O/${TAG}_hop.txt 2>&1
    cat $O/${TAG}_hop.txt ;;
hopab)
    # same box, same minute: the round-1 library (build/r1, made from commit 315125a) against the current one
    : > $O/${TAG}_hopab.txt
    for i in 1 2; do
        for sz in 512x512x512 256x256x256; do
            (cd build/r1 && timeout 200 python scripts/hop_bench.py 2 $sz 2>&1 | sed 's/^/r1  /') >> $O/${TAG}_hopab.txt
            timeout 200 python scripts/hop_bench.py 2 $sz 2>&1 | sed 's/^/now /' >> $O/${TAG}_hopab.txt
            MGCR_PDL=0 timeout 200 python scripts/hop_bench.py 2 $sz 2>&1 | sed 's/^/now pdl=0 /' >> $O/${TAG}_hopab.txt
        done
    done
    cat $O/${TAG}_hopab.txt ;;
dist)
    # multi-GPU parity + bench at N = $NGPU (gpurun --gpus N): tests for this world size, then the bench with the deferred halo wait on / off
    N=${NGPU:-2}
    timeout 1500 python -m pytest tests/test_gpu_dist.py -m gpu -q -k "test_distributed_against_single_gpu and ${N}-" > $O/${TAG}_pytest_dist${N}.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_dist${N}.log
    grep -E "passed|failed|skipped|rc=" $O/${TAG}_pytest_dist${N}.log | tail -4
    grep -E "^E  " $O/${TAG}_pytest_dist${N}.log | cut -c1-400 | head -20
    first=1
    for cfg in ${DIST_CFGS-"MGCR_HALO_DEFER=1" "MGCR_HALO_DEFER=0"}; do
        extra="--no-others"; [ $first = 1 ] && extra=""; first=0     # the first configuration is the driver's command (secondary workloads included)
        env $cfg timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus $N --steps 3 --warmup 2 --no-cpu-baseline $extra \
            > $O/${TAG}_bench_n${N}_${cfg//=/}.json 2> $O/${TAG}_bench_n${N}_${cfg//=/}.err; echo "== $cfg rc=$?"
        python scripts/bench_brief.py $O/${TAG}_bench_n${N}_${cfg//=/}.json 2>&1 | head -24
    done ;;
q256)
    # the 1/8-size problem (the per-GPU sizes of the 8-GPU run) on one GPU, virtual slabs on / off
    : > $O/${TAG}_q256.txt
    for cfg in "MGCR_RED_VSLABS=8" "MGCR_RED_VSLABS=1"; do
        echo "== $cfg" >> $O/${TAG}_q256.txt
        env $cfg timeout 200 python bench.py --workload mg3d_256 --operator stencil --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 > $O/knob_tmp.json
        python scripts/bench_brief.py $O/knob_tmp.json 2>/dev/null | head -13 >> $O/${TAG}_q256.txt
    done
    cat $O/${TAG}_q256.txt ;;
small)
    # persistent small-system solver: CTAs per SM x rows per CTA (the replicated coarse solves are 9 % of the 8-GPU solve)
    : > $O/${TAG}_small_sweep.txt
    for ps in 1 2 3; do for rows in 128 256 512 1024; do
        echo "== per_sm=$ps rows_per_cta=$rows" >> $O/${TAG}_small_sweep.txt
        MGCR_SMALL_GRID_PER_SM=$ps MGCR_SMALL_ROWS_PER_CTA=$rows timeout 200 python bench.py --workload mg3d_256 --operator stencil --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 > $O/knob_tmp.json
        python scripts/bench_brief.py $O/knob_tmp.json 2>/dev/null | grep -E "^N=|gcr_small|blockcsr" >> $O/${TAG}_small_sweep.txt
    done; done
    cat $O/${TAG}_small_sweep.txt ;;
knobs)
    # same box: programmatic dependent launch and blind preconditioned solves on / off
    : > $O/${TAG}_knobs.txt
    for cfg in ${KNOB_CFGS:-"MGCR_PDL=1 MGCR_BLIND_PRECOND=1" "MGCR_PDL=0 MGCR_BLIND_PRECOND=1" "MGCR_PDL=1 MGCR_BLIND_PRECOND=0" "MGCR_PDL=0 MGCR_BLIND_PRECOND=0" "MGCR_PDL=1 MGCR_BLIND_PRECOND=1"}; do
        echo "== $cfg" >> $O/${TAG}_knobs.txt
        env $cfg timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-others 2>/dev/null | tail -1 > $O/knob_tmp.json
        python scripts/bench_brief.py $O/knob_tmp.json 2>/dev/null | head -${KNOB_LINES:-2} >> $O/${TAG}_knobs.txt
    done
    cat $O/${TAG}_knobs.txt ;;
ab)
    # same box, interleaved: the library of an earlier commit (build/prev: git archive <commit> | tar -x -C build/prev; make) against the current one
    : > $O/${TAG}_ab.txt
    for i in 1 2 3; do
        for which in prev now; do
            d=.; [ $which = prev ] && d=build/prev
            (cd $d && timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-others 2>/dev/null | tail -1) > $O/knob_tmp.json
            echo "== $which" >> $O/${TAG}_ab.txt
            python scripts/bench_brief.py $O/knob_tmp.json 2>/dev/null | head -${KNOB_LINES:-13} | grep -v "^mg_setup\|^roofline" >> $O/${TAG}_ab.txt
        done
    done
    cat $O/${TAG}_ab.txt ;;
kbt)
    # standalone layout comparison for restrict / prolong (scripts/kbench_transfer.cu, built into build/kbt)
    for n in 256 512; do timeout 120 build/kbt $n; done > $O/${TAG}_kbench_transfer.txt 2>&1; cat $O/${TAG}_kbench_transfer.txt ;;
bench)
    timeout 900 python bench.py > $O/${TAG}_bench_default_n1.json 2> $O/${TAG}_bench_default_n1.err; echo "bench rc=$?"
    python scripts/bench_brief.py $O/${TAG}_bench_default_n1.json ;;
quick)
    timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-others > $O/${TAG}_bench_quick_n1.json 2> $O/${TAG}_bench_quick_n1.err; echo "quick rc=$?"
    python scripts/bench_brief.py $O/${TAG}_bench_quick_n1.json ;;
launches)
    CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-others --max-iter 2"
    $CMD > $O/ncu_plain.log 2>&1 || { echo plain failed; tail -5 $O/ncu_plain.log; continue; }
    timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file $O/${TAG}_launches_mg3d_512.csv $CMD > $O/ncu_a.log 2>&1; tail -1 $O/ncu_a.log ;;
ncufull)
    CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-others --max-iter 2"
    $CMD > $O/ncu_plain.log 2>&1 || { echo plain failed; tail -5 $O/ncu_plain.log; continue; }
    # set-up launches come first (Arnoldi: ~140 stencil applies); skip them so that the captures are the solve's fine-level launches
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_restrict_warp|k_prolong' --launch-skip 0 -c 6 -f -o $O/${TAG}_mg512_transfer $CMD > $O/ncu_c.log 2>&1; tail -1 $O/ncu_c.log
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_hopping_tma' --launch-skip 150 -c 8 -f -o $O/${TAG}_mg512_stencil $CMD > $O/ncu_d.log 2>&1; tail -1 $O/ncu_d.log
    ls -la $O/*.ncu-rep ;;
*) echo "unknown stage $stage" ;;
esac
done
