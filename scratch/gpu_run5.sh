mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for c in hyperbolic_test elliptic_test; do echo "== $c"; ./impec2d_b200/impec2d examples/$c.cfg 2>&1 | tail -8; done
echo "== parallel.cfg 30 steps"; ./impec2d_b200/impec2d examples/parallel.cfg --steps 30 2>&1 | tail -6
echo "== diagonal_dispersion.cfg 10 steps"; ./impec2d_b200/impec2d examples/diagonal_dispersion.cfg --steps 10 2>&1 | tail -5
echo "== parallel_dispersion.cfg 10 steps"; ./impec2d_b200/impec2d examples/parallel_dispersion.cfg --steps 10 2>&1 | tail -5
cat /tmp/hyperbolic_norms.out /tmp/elliptic_norms.out
