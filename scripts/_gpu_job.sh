cd /root/repo
cat > /tmp/tiny_fwd.py <<'PY'
import sys, torch
for p in ("/root/repo", "/root/repo/vla-from-fastvlm_b200", "/root/repo/tests", "/root/repo/tests/golden"):
    sys.path.insert(0, p)
from helpers import TINY_HEAD, make_engine, make_inputs, tiny_weights
arch, sd, hsd = tiny_weights(0)
eng = make_engine(arch, sd, hsd, torch.bfloat16)
for B in (1, 5):
    images, states, ids, mask = make_inputs(B, 120, 160, 9, arch.text.vocab, TINY_HEAD["state_dim"], seed=3, image_mode="prefix")
    for _ in range(3):
        out = eng.forward(images.to(eng.device), ids, mask.sum(1), states=states.to(eng.device))
    torch.cuda.synchronize()
    print(B, out.float().abs().mean().item())
PY
echo "== memcheck"; timeout 400 compute-sanitizer --tool memcheck --error-exitcode 3 python /tmp/tiny_fwd.py 2>&1 | tail -6
echo "== synccheck"; timeout 300 compute-sanitizer --tool synccheck --error-exitcode 3 python /tmp/tiny_fwd.py 2>&1 | tail -4
