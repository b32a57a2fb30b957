"""Pretty-print the event timeline a -DFVLA_FFN_TRACE_BUILD libfvla.so writes on stderr (scripts/time_ffn.py 2> log)."""
import collections, sys
ev = []
for l in open(sys.argv[1]):
    if l.startswith('TR'):
        _, r, i, t = l.split(); ev.append((int(t), int(r), int(i)))
ev.sort()
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 200)
for t, r, i in ev[lo:hi]:
    if r == 0:
        base = (i // 100) * 100; k = i - base
        d = {1000: 'G1 begin', 1100: 'G1 S free', 1200: 'G1 W1+X ok', 1300: 'G1 issued'}[base] + f' chunk {k}'
        who = 'G1W '
    elif r == 1:
        base = (i // 200) * 200; k = i - base
        d = {2000: 'G2 begin', 2200: 'G2 H ok', 2400: 'G2 issued'}[base] + f' chunk {k // 2} half {k % 2}'
        who = ' G2W'
    else:
        base = (i // 100) * 100; k = i - base
        d = {100: 'wait S', 200: 'got S', 300: 'math done', 400: 'got hempty', 500: 'stored+arrived'}[base] + f' chunk {k}'
        who = f'   EPI{r - 2}'
    print(f'{t:8d} {who:8s} {d}')
