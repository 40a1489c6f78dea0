#!/bin/bash
O=gpurun_out/r2j; mkdir -p $O
( for n in 2 4 8; do timeout 120 python tools/gang_probe.py $n 12 sp; done
timeout 120 python tools/gang_probe.py 4 12 dp ) > $O/probe.log 2>&1
cat $O/probe.log | cut -c1-300 | tail -20
timeout 1200 python -m pytest tests/test_gang.py tests/test_lammps_deck.py -m gpu -x -q > $O/pytest.log 2>&1; echo "exit $?" >> $O/pytest.log
tail -30 $O/pytest.log
