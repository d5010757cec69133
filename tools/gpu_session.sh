#!/bin/sh
# One parametrised entry for every gpurun call of the round (see tools/experiments/README.md).
# usage: sh tools/gpu_session.sh <tag> <stage> [stage args] [-- <stage> ...]
# Stages run in order; a failing stage does not stop the later ones (each writes its own log).
TAG="$1"; shift
OUT=gpurun_out
mkdir -p $OUT
run_stage() {
	stage="$1"; shift
	echo "=== stage $stage $* ($(date +%T))"
	case "$stage" in
	tests)
		timeout 1200 python -m pytest tests -m gpu -x -q "$@" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/${TAG}_pytest.log ;;
	bench)
		name="$1"; shift
		timeout 1500 python bench.py "$@" > $OUT/${TAG}_bench_${name}.json 2> $OUT/${TAG}_bench_${name}.err; echo "bench exit $?"
		tail -c 600 $OUT/${TAG}_bench_${name}.err; head -c 1500 $OUT/${TAG}_bench_${name}.json; echo ;;
	ab)
		for v in 1 0; do RTK_B200_L2_PERSIST=$v timeout 300 python tools/prof_trace.py C3 6 2>&1 | tail -1 | sed "s/^/L2_PERSIST=$v /"; done
		for v in 1 0; do RTK_B200_L2_PERSIST=$v timeout 600 python tools/prof_trace.py C4 4 33554432 2>&1 | tail -1 | sed "s/^/L2_PERSIST=$v /"; done
		for v in 1 0; do RTK_B200_HOST_DIRECT=$v timeout 600 python tools/prof_e2e.py C3 16777216 1 5 2>&1 | tail -12 | sed "s/^/HOST_DIRECT=$v /"; done ;;
	ncu)
		w="$1"; n="${2:-16777216}"
		timeout 600 python tools/prof_trace.py $w 4 $n > $OUT/${TAG}_prof_${w}_plain.log 2>&1 &&
		timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 2 -c 1 -f -o $OUT/${TAG}_k_trace_${w} \
			python tools/prof_trace.py $w 4 $n > $OUT/${TAG}_prof_${w}_ncu.log 2>&1
		echo "ncu exit $?"; tail -2 $OUT/${TAG}_prof_${w}_plain.log ;;
	launches)
		timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --parity-rays 0 --legs e2e > $OUT/${TAG}_launch_plain.log 2>&1 &&
		timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/${TAG}_launches_bench.csv \
			python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --parity-rays 0 --legs e2e > $OUT/${TAG}_launch_ncu.log 2>&1
		echo "launch list exit $?" ;;
	build)
		for w in C3 C4 C2; do timeout 600 python tools/build_profile.py --config $w --rebuilds 3 2>&1 | tail -3 | sed "s/^/$w /"; done ;;
	buildncu)
		timeout 600 python tools/prof_build.py ${1:-C3} sah > $OUT/${TAG}_pb_plain.log 2>&1 &&
		timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/${TAG}_launches_build_${1:-C3}.csv \
			python tools/prof_build.py ${1:-C3} sah > $OUT/${TAG}_pb_ncu.log 2>&1
		echo "build launch list exit $?"; tail -1 $OUT/${TAG}_pb_plain.log ;;
	scale)
		# the N-rank line with the NCCL gather (all legs) and with the peer-memory gather (headline only)
		N="$1"; shift
		for g in nccl p2p; do
			legs="--legs c4,e2e"; [ "$g" = p2p ] && legs="--legs none --parity-rays 0"
			timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
				bench.py --gpus $N --steps 30 --warmup 3 --gather $g $legs "$@" > $OUT/${TAG}_scale_${N}_${g}.json 2> $OUT/${TAG}_scale_${N}_${g}.err
			echo "scale $N $g exit $?"; tail -c 400 $OUT/${TAG}_scale_${N}_${g}.err; head -c 600 $OUT/${TAG}_scale_${N}_${g}.json; echo
		done ;;
	scale8final|benchn)
		# the driver's own command line at N ranks: bench.py defaults (peer-memory gather, all legs)
		N="${1:-8}"
		timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
			bench.py --gpus $N --steps 30 --warmup 3 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
		echo "bench $N exit $?"; tail -c 300 $OUT/${TAG}_bench_n$N.err; head -c 400 $OUT/${TAG}_bench_n$N.json; echo ;;
	sanitize)
		# compute-sanitizer memcheck over the GPU cases that exercise the kernels added in round 2 (one tool per call)
		timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -x -q \
			-k "known_answer or direct_rows or blob_validation or two_streams or probes or configs_reduced or refit" > $OUT/${TAG}_sanitizer.log 2>&1
		echo "sanitizer exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|error" $OUT/${TAG}_sanitizer.log | tail -8 ;;
	e2e)
		# one process, N devices: rows direct / staged, compact, link ceilings
		N="$1"
		timeout 900 python tools/prof_e2e.py C3 $((16777216 * N)) $N 4 2>&1 | tail -14
		RTK_B200_HOST_DIRECT=0 timeout 900 python tools/prof_e2e.py C3 $((16777216 * N)) $N 4 2>&1 | grep -E "rows " | sed "s/^/HOST_DIRECT=0 /" ;;
	variants)
		# k_trace of every variant library rtk_b200/librtk_b200_<name>.so on one workload: "variants C3 base v00 v10 ..."
		w="$1"; shift
		for v in "$@"; do
			lib=rtk_b200/librtk_b200_$v.so; [ "$v" = main ] && lib=rtk_b200/librtk_b200.so
			RTK_LIB=$lib timeout 300 python tools/prof_trace.py $w 6 2>&1 | tail -1 | sed "s/^/$v /"
		done ;;
	mix)
		for m in 0 2 3 4; do RTK_B200_HOST_MIX=$m timeout 600 python tools/prof_e2e.py C3 16777216 1 5 2>&1 | grep -E "rows  " | sed "s/^/MIX=$m /"; done
		for r in 6 12; do RTK_B200_PUSH_SMS=$r RTK_B200_HOST_MIX=2 timeout 600 python tools/prof_e2e.py C3 16777216 1 5 2>&1 | grep -E "rows  " | sed "s/^/MIX=2 PUSH_SMS=$r /"; done ;;
	buildvariants)
		# device build time of variant libraries: "buildvariants main hyb ..."
		for v in "$@"; do
			lib=rtk_b200/librtk_b200_$v.so; [ "$v" = main ] && lib=rtk_b200/librtk_b200.so
			for w in C3 C4 C2; do RTK_LIB=$lib timeout 300 python tools/prof_build.py $w sah 2>&1 | tail -1 | sed "s/^/$v /"; done
		done ;;
	sortbits)
		# Morton bits the SAH builder's input is sorted by (one radix pass per 8): device build time
		for b in 32 24 16; do for w in C3 C4; do RTK_B200_SAH_SORT_BITS=$b timeout 300 python tools/prof_build.py $w sah 2>&1 | tail -1 | sed "s/^/SORT_BITS=$b /"; done; done ;;
	buildfull)
		# ncu --set full of the build's heavy kernels (first build of the run)
		timeout 600 python tools/prof_build.py ${1:-C3} sah > $OUT/${TAG}_pbf_plain.log 2>&1 &&
		timeout 1200 ncu --set full --clock-control none -k regex:'k_sah_small|k_sah_bin_large|k_collapse|k_sah_partition_large|k_radix_scatter' -c 40 -f -o $OUT/${TAG}_build_${1:-C3} \
			python tools/prof_build.py ${1:-C3} sah > $OUT/${TAG}_pbf_ncu.log 2>&1
		echo "build ncu exit $?"; tail -1 $OUT/${TAG}_pbf_plain.log ;;
	hostab)
		for v in 1 0; do RTK_B200_HOST_DIRECT=$v timeout 600 python tools/prof_e2e.py C3 16777216 1 5 2>&1 | grep -E "rows|compact" | sed "s/^/HOST_DIRECT=$v /"; done
		for r in 2 8; do RTK_B200_PUSH_SMS=$r timeout 600 python tools/prof_e2e.py C3 16777216 1 5 2>&1 | grep -E "rows " | sed "s/^/PUSH_SMS=$r /"; done ;;
	topo)
		# what the host looks like: sockets, NUMA nodes, which node each GPU hangs off, the cpuset of this container
		lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)|Thread|Core" ; nproc
		for n in /sys/devices/system/node/node*; do echo "$n: cpus $(cat $n/cpulist) $(grep MemTotal $n/meminfo | tr -s ' ')"; done
		for d in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader | tr 'A-Z' 'a-z' | sed 's/^0000//'); do echo "gpu $d numa_node $(cat /sys/bus/pci/devices/$d/numa_node 2>/dev/null)"; done
		nvidia-smi topo -m 2>&1 | head -14
		grep -E "Cpus_allowed_list|Mems_allowed_list" /proc/self/status ;;
	numa)
		# host link ceilings and the row path under the three placement policies of the batch arrays
		N="$1"
		for v in 1 0; do RTK_B200_NUMA=$v timeout 900 python tools/prof_e2e.py C3 $((16777216 * N)) $N 3 2>&1 | grep -E "rows  |compact|host link, $N" | sed "s/^/NUMA=$v /"; done
		RTK_B200_HOST_MIX=0 timeout 900 python tools/prof_e2e.py C3 $((16777216 * N)) $N 3 2>&1 | grep -E "rows  " | sed "s/^/MIX=0 /" ;;
	*) echo "unknown stage $stage" ;;
	esac
}
args=""
while [ $# -gt 0 ]; do
	if [ "$1" = "--" ]; then run_stage $args; args=""; else args="$args $1"; fi
	shift
done
[ -n "$args" ] && run_stage $args
echo "=== done ($(date +%T))"
