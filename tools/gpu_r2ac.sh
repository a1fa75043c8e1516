# Round 2, call AC: the timeline of back-to-back calls in the stream (globaltimer stamps from inside the kernels)
set -x
mkdir -p gpurun_out
{
./tools/timeline_probe 150 40 32 1000 8
./tools/timeline_probe 400 80 8 1024 6
} > gpurun_out/r2ac_timeline.txt 2>&1
cat gpurun_out/r2ac_timeline.txt
