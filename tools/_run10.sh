make -s -C examples && ./examples/_build/covariance_check; ./examples/_build/robot_3d_localization | tail -7
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
