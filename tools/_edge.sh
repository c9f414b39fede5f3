for e in 0 1 4 20; do echo "== EB2_EDGE_CHUNKS=$e"; EB2_EDGE_CHUNKS=$e python tools/exp_lane.py 0 160 2>&1 | tail -9; done
python tools/quick_parity.py --no-timing 2>&1 | tail -3
