# 1 GPU: the full GPU parity suite on every fallback route (one line per route)
for env in "" "B200_NO_SPEC=1" "B200_SNAKE=0" "B200_NBODY_FUSED=2" "B200_NBODY_WORLD_MIN=444 B200_NBODY_WORLD_ROUNDS=2 B200_NBODY_WORLD_SPREAD=16" "B200_SMALL_WORLD=0" "B200_EXACT_CFG=12" "B200_NBODY_FUSED=0 B200_GRAPH_CFG=0" "B200_CHUNK_BODIES=4096"; do
  r=$(env $env timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -1)
  printf "%-90s %s\n" "${env:-default routes}" "$r"
done
