# Round 2, call W: the zero fill alone -- run list with the pieces of the compacted byte space interleaved over the grid,
# bulk stores and 16-byte stores, on the scattered runs and on ONE run of the same size
set -x
mkdir -p gpurun_out
{
./tools/fill_probe 150 40 32 1000
./tools/fill_probe 400 80 32 1024
} > gpurun_out/r2w_fill_probe.txt 2>&1
cat gpurun_out/r2w_fill_probe.txt
