#!/bin/bash
# one GPU pass: parity tests, C2 bench, launch list, ncu captures of the dominant kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2}
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_pytest.log | tail -5
python bench.py --steps 200 --warmup 10 > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err
cut -c1-1500 gpurun_out/${TAG}_bench_c2.json
if [ "$2" != "noncu" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu --no-extra > gpurun_out/${TAG}_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_raster_quads -s 40 -c 1 -f -o gpurun_out/${TAG}_prof_k1 \
    python bench.py --steps 8 --warmup 3 --no-cpu --no-extra > gpurun_out/${TAG}_ncu_k1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_render_color -s 40 -c 1 -f -o gpurun_out/${TAG}_prof_k4 \
    python bench.py --steps 8 --warmup 3 --no-cpu --no-extra > gpurun_out/${TAG}_ncu_k4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_frame_prologue -s 40 -c 1 -f -o gpurun_out/${TAG}_prof_k0 \
    python bench.py --steps 8 --warmup 3 --no-cpu --no-extra > gpurun_out/${TAG}_ncu_k0.log 2>&1
# (three 19 MB reports exceed what gpurun copies back: keep the text summaries only)
for k in k1:k_raster_quads k4:k_render_color k0:k_frame_prologue; do
  python tools/ncu_summary.py gpurun_out/${TAG}_prof_${k%%:*}.ncu-rep > gpurun_out/${TAG}_${k##*:}_ncu.txt 2>/dev/null
  rm -f gpurun_out/${TAG}_prof_${k%%:*}.ncu-rep
done
fi
ls -la gpurun_out | tail -8
