# ncu evidence for the round-2 build.  Run on one B200:  gpurun -- 'bash profiles/ncu_r2.sh'
# Every program first runs to completion WITHOUT ncu; numbers printed under ncu are never bench values.
#  1. launch list of the driver's bench command shape (3 warm-up + 2 timed steps + stage pass + e2e)
#  2. full capture of the 12 hot-path launches of the first timed step  -> profiles/r2_ncu_step_kernels.json, traffic.json
#  3. full captures of the K4 inverse ring kernel and of the streamed-bank CQT contraction (whole transform + frame window)
set -x
B="python bench.py --steps 2 --warmup 3 --cpu-windows 0 --e2e-steps 3 --no-extras --no-ref-shape"
K='regex:stft_ring|stft_kernel|cqt_|decimate|subtract_|window_db'
FULL="ncu --set full --clock-control none --import-source on"
$B > gpurun_out/r2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 150 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_ncu_a.log 2>&1
$FULL -k "$K" -s 36 -c 12 -f -o gpurun_out/r2_step $B > gpurun_out/r2_ncu_b.log 2>&1
python profiles/microbench/istft_k4_only.py > gpurun_out/r2_k4_plain.log 2>&1 && \
$FULL -k regex:istft_ring -s 2 -c 1 -f -o gpurun_out/r2_k4 python profiles/microbench/istft_k4_only.py > gpurun_out/r2_ncu_c.log 2>&1
python profiles/microbench/cqt_stream_only.py > gpurun_out/r2_stream_plain.log 2>&1 && \
$FULL -k regex:cqt_umma_stream -s 1 -c 4 -f -o gpurun_out/r2_stream python profiles/microbench/cqt_stream_only.py > gpurun_out/r2_ncu_d.log 2>&1
tail -3 gpurun_out/r2_plain.log gpurun_out/r2_k4_plain.log gpurun_out/r2_stream_plain.log
ls -la gpurun_out/ | tail -12
