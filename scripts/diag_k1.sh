for a in "96 128 1 bf16 bias" "96 128 1 bf16" "96 96 1 bf16" "128 128 1 bf16" "96 128 1 f32 bias" "128 32 1 f32"; do
  python scripts/conv_diag.py floor2 5 $a 2>&1 | tail -1
  MK3D_TRACE=1 python scripts/conv_diag.py floor2 2 $a 2>&1 | grep "mk3d trace" | tail -1
done
