timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --prefetch 1 2>/dev/null | python tools/brief.py prefetch
