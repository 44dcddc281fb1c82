# ncu evidence for the end-of-round build (v13: serial schedule, every cascade pair fused, flat single-step subtract,
# deep-load dB kernel, PCM ingest): launch list of all hot kernels (3 warm-up + 2 timed steps + stage pass + e2e), then
# full captures of the 12 launches of the first timed step and of the K0 kernels.  Run: gpurun -- 'bash profiles/ncu_r1_v13.sh'
set -x
B="python bench.py --steps 2 --warmup 3 --cpu-windows 0 --e2e-steps 3"
K='regex:stft_kernel|cqt_|decimate|subtract_|window_db|pcm16'
$B > gpurun_out/plain13.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 120 --csv --log-file gpurun_out/launches_v13.csv $B > gpurun_out/ncu13a.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:stft_kernel|cqt_|decimate|subtract_|window_db' -s 36 -c 12 -f -o gpurun_out/prof_v13 $B > gpurun_out/ncu13b.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:pcm16' -s 8 -c 4 -f -o gpurun_out/prof_v13_pcm $B > gpurun_out/ncu13c.log 2>&1
ls -la gpurun_out/
