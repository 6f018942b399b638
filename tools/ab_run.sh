for i in 1 2; do
for v in "" build_variants/lib_head.so; do echo "== $v"; TINA_B200_LIB=$v python tools/ab_knob.py pdl 1 2>&1 | tail -1; done
done
