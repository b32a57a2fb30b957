timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "a_stationary or test_gemm_bf16 or fp16_hidden" 2>&1 | tail -5
for i in 1 2; do
python scripts/profile_forward.py --batch 64 --steps 6 --warmup 3 2>&1 | grep -E "^# forward|N1536 K384 gelu" | cut -c1-110
FVLA_DISABLE_GEMM_AST=1 python scripts/profile_forward.py --batch 64 --steps 6 --warmup 3 2>&1 | grep -E "^# forward|N1536 K384 gelu" | cut -c1-110
done
