#!/bin/bash
# round 2, pass I (1 GPU): gang handle (several bricks in one process), decks on several bricks, thermo without download
O=gpurun_out/r2i; mkdir -p $O
timeout 1200 python -m pytest tests/test_gang.py tests/test_lammps_deck.py -m gpu -x -q > $O/pytest.log 2>&1; echo "exit $?" >> $O/pytest.log
tail -30 $O/pytest.log
