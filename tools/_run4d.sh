for m in 0 1 2; do timeout 200 python tools/debug_multi.py 4 $m 2>&1 | grep "^mode" | sort; done
timeout 200 python tools/debug_multi.py 3 1 2>&1 | grep "^mode" | sort
timeout 200 python tools/debug_multi.py 4 1 120000 640 384 2>&1 | grep "^mode" | sort
