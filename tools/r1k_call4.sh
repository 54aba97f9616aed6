#!/bin/bash
# GPU call 4 (r1k): smoke() and the GPU parity suite on the final tree of the round
out=gpurun_out; mkdir -p $out
( time timeout 40 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > $out/r1k_smoke.log 2>&1; tail -4 $out/r1k_smoke.log
( time timeout 150 python -m pytest tests -m gpu -x -q ) > $out/r1k_pytest.log 2>&1; tail -6 $out/r1k_pytest.log
