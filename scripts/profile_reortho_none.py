"""`tridiag(reortho="none")` at the published benchmark's shape (bench_extra.reortho_none): forward + adjoint once more after a warm-up."""
import sys

sys.path.insert(0, ".")
import bench_extra  # noqa: E402
import experiments_lanczos_adjoints_b200 as bl  # noqa: E402

print(bench_extra.reortho_none(bl, reps=2))
