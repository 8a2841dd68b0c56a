for st in 0 1000 2000 3000 4000; do echo "stagger $st"; RB_ROLLOUT_STAGGER_NS=$st python bench.py --steps 50 --warmup 5 --quick 2>&1 | grep -o '"value": [0-9.]*' | head -1; done
