set -x
mkdir -p gpurun_out
{
echo "== 204.8 MB (c2's zero rows)"; ./tools/zero_probe 51200 1000 | head -24
echo "== 1.12 GB (c3's zero rows)"; ./tools/zero_probe 273000 1024 | head -24
} > gpurun_out/r2t_zero_burst.txt 2>&1
cat gpurun_out/r2t_zero_burst.txt
