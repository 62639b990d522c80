make -s -C examples && ./examples/_build/bundle_adjustment
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
./examples/_build/bundle_adjustment_in_the_large --synthetic 1778 993923 5 | grep -i "timing\|Created\|Finished"
