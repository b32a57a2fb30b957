for h in 0 3 7 2 0 3; do
echo "hints $h"; FVLA_L2_HINTS=$h python scripts/profile_forward.py --batch 64 --steps 6 --warmup 3 2>&1 | grep -E "^# forward|N1536 K384 gelu|N384 K1536|k3 s1 m1 C384|k7 s1 m1 C384" | cut -c1-110
done
