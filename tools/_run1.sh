set -x
mkdir -p gpurun_out
for v in base u5 u6 u8; do
echo "== $v" >> gpurun_out/r02l_unproj.log
WR_B200_LIB=$PWD/worldrenderer_b200/lib/variants/lib_$v.so timeout 300 python tools/bake_graph_probe.py >> gpurun_out/r02l_unproj.log 2>&1
WR_B200_LIB=$PWD/worldrenderer_b200/lib/variants/lib_$v.so timeout 300 python tools/bake_phases.py 2>&1 | tail -4 >> gpurun_out/r02l_unproj.log
done
cat gpurun_out/r02l_unproj.log
