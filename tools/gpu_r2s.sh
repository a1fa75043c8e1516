set -x
mkdir -p gpurun_out
{
echo "== k2_probe T=400 S=80 B=8 parts=8 K=1 zero=0 V=1024"; ./tools/k2_probe 400 80 8 8 1 0 1024
echo "== k2_probe T=400 S=80 B=8 parts=8 K=1 zero=2 V=1024"; ./tools/k2_probe 400 80 8 8 1 2 1024
echo "== k2_probe T=400 S=30 B=8 parts=8 K=1 zero=0 V=1024"; ./tools/k2_probe 400 30 8 8 1 0 1024
echo "== k2_probe T=150 S=40 B=32 parts=4 K=1 zero=0 V=1000"; ./tools/k2_probe 150 40 32 4 1 0 1000
echo "== k2_probe T=150 S=40 B=32 parts=4 K=1 zero=2 V=1000"; ./tools/k2_probe 150 40 32 4 1 2 1000
} > gpurun_out/r2s_probe.txt 2>&1
cat gpurun_out/r2s_probe.txt | cut -c1-1200
