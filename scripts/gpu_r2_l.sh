# round 2, session 2, call 9: full GPU suite incl. the INTEGRATION.md binding test; smoke
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2l_tests.log 2>&1
echo "pytest exit $?" >> $O/r2l_tests.log
tail -12 $O/r2l_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
