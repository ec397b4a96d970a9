for d in 0 4; do echo "== AISB_FAST_DEBUG=$d"; AISB_FAST_DEBUG=$d python scripts/kde_screen_check.py 0.25 1 2>&1 | grep -A4 "fast - plain\|fast units"; done
python scripts/kde_screen_check.py 1.0 3 2>&1 | tail -22
