python -m pytest tests -m gpu -x -q 2>&1 | tail -n 2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
