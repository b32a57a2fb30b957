L=vla-from-fastvlm_b200/vla_fastvlm/_lib/libfvla.so
cp $L /tmp/base.so
for i in 1 2; do
cp /tmp/base.so $L
echo base; for b in 1 64; do python scripts/profile_forward.py --batch $b --steps 8 --warmup 3 2>&1 | grep "^# forward"; done
cp vla-from-fastvlm_b200/csrc/build_early/libfvla_early.so $L
echo early; for b in 1 64; do python scripts/profile_forward.py --batch $b --steps 8 --warmup 3 2>&1 | grep "^# forward"; done
done
timeout 300 python -m pytest tests -m gpu -x -q -k "pdl or engine" 2>&1 | tail -2
