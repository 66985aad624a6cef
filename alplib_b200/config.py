"""Package-wide defaults."""

# "fast": algebraically simplified per-sample arithmetic, <= 1e-9 relative from the reference (the parity tolerance).
# "faithful": the reference's operation order expression by expression.
PRECISION = "fast"
