echo "== graph replay at the bench shapes"; timeout 900 python tools/diag_graph_shapes.py 2>&1 | tail -5
echo "== pytest gpu full"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
