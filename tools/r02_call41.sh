set -x
for lib in build/libphylo_nw13.so build/libphylo_nw7x3.so; do
  PHYLO_B200_LIB=$PWD/$lib python tools/sweep_dense_cherry.py 2>&1 | grep "^{" | grep '"c3"' | cut -c1-200 | head -2 | sed "s|^|$lib |"
done | tee gpurun_out/r03c_sweep_dense_shapes.txt
