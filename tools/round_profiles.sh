#!/bin/bash
# Everything the round's profiles/ are made from, in one gpurun call (one B200):
#   gpurun --timeout 1500 -- 'bash tools/round_profiles.sh r02'
# Writes gpurun_out/<tag>_*.{json,log,csv}; ncu reports stay in /tmp on the box, only their CSV exports come back.
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1

if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_gputests.log 2>&1; echo "gpu tests rc=$?" >> $out/${tag}_gputests.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/${tag}_smoke.log 2>&1
fi

# bench lines (config 2 with its CPU leg; the others without)
timeout 600 python bench.py > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err
timeout 300 python bench.py --config 1 --no-cpu-baseline > $out/${tag}_bench_c1.json 2> $out/${tag}_bench_c1.err
timeout 300 python bench.py --config 3 --no-cpu-baseline --steps 50 > $out/${tag}_bench_c3_ddim10.json 2> $out/${tag}_bench_c3_ddim10.err
timeout 300 python bench.py --config 3 --ddim-steps 50 --no-cpu-baseline --steps 20 > $out/${tag}_bench_c3_ddim50.json 2> $out/${tag}_bench_c3_ddim50.err
timeout 300 python bench.py --config 4 --no-cpu-baseline --steps 50 > $out/${tag}_bench_c4.json 2> $out/${tag}_bench_c4.err
timeout 300 python bench.py --config 5 --no-cpu-baseline --steps 20 > $out/${tag}_bench_c5.json 2> $out/${tag}_bench_c5.err

# launch list of the bench command (shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_launches.log 2>&1

# full capture of the sampler kernel (10 DDIM steps of config 2's 1280 samples: same per-step behaviour, short replay)
GLDM_TC_ROWS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:resnet_rows -c 1 -f -o /tmp/${tag}_rows \
  python tools/run_sampler_once.py 100 1280 bf16 > $out/${tag}_ncu_rows.log 2>&1
ncu -i /tmp/${tag}_rows.ncu-rep --page raw --csv > $out/${tag}_rows_raw.csv 2>/dev/null
ncu -i /tmp/${tag}_rows.ncu-rep --page source --csv > $out/${tag}_rows_src.csv 2>/dev/null

# full capture of one encoder pass (64 clouds)
GLDM_PROFILE_RANGE=1 timeout 900 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/${tag}_enc \
  python tools/run_encoder_once.py 64 bf16 3 1 > $out/${tag}_ncu_enc.log 2>&1
ncu -i /tmp/${tag}_enc.ncu-rep --page raw --csv > $out/${tag}_enc_raw.csv 2>/dev/null


# operator FFI against the reference's own kernels (oracle/_ref), per-row-job stamps of the sampler and of the decoder
timeout 600 python tests/tools/bench_ops.py > $out/${tag}_ops_bench.log 2>&1 && cp $out/ops_bench.json $out/${tag}_ops_bench.json
timeout 200 python tools/prof_rows.py sampler > $out/${tag}_prof_rows.log 2>&1
timeout 200 python tools/prof_rows.py decoder > $out/${tag}_prof_rows_decoder.log 2>&1

ls -la $out | tail -30
