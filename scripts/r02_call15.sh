cd /root/repo
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python scripts/gpu_big_parity.py 2>&1 | tail -6
