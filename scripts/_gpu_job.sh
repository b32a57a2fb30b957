python scripts/profile_forward.py --batch 1 --steps 8 --warmup 3 2>&1 | head -3 | tail -1
python scripts/profile_forward.py --batch 2 --steps 8 --warmup 3 2>&1 | head -3 | tail -1
timeout 600 python -m pytest tests -m gpu -x -q -k "engine or fullsize or decoder_shapes or golden or plugin" 2>&1 | tail -3
