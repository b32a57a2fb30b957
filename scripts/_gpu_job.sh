for args in "2 64 96 7 1 1" "2 64 96 7 2 2" "2 64 96 3 1 1"; do
  echo "== memcheck $args"; compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/one_dwconv.py $args 2>&1 | tail -3
done
echo "== racecheck 7 1 1"; timeout 300 compute-sanitizer --tool racecheck --error-exitcode 3 python scripts/one_dwconv.py 1 32 32 7 1 1 2>&1 | tail -4
echo "== racecheck 7 2 2"; timeout 300 compute-sanitizer --tool racecheck --error-exitcode 3 python scripts/one_dwconv.py 1 32 32 7 2 2 2>&1 | tail -4
