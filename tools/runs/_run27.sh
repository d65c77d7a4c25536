mkdir -p gpurun_out
for shp in "16 512 128 3 1" "16 256 200 3 1"; do
  tag=$(echo $shp | tr ' ' '_')
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"tc_conv|tc_wgrad" -s 6 -c 3 -o gpurun_out/r02_small_$tag python tools/profile_convblock.py tc $shp > gpurun_out/ncu_small_$tag.log 2>&1; tail -1 gpurun_out/ncu_small_$tag.log
done
