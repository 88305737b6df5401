"""Numerical constants of the network-crossover path (values from the reference's
`parameters.py:7-8,16`; the perturbation-crossover constants at :22-28 belong to a
different algorithm and are out of scope)."""

# an artificial (big-M) arc counts as "out of the solution" below this flow
TOLERANCE_FOR_ARTIFICIAL_VARS = 1e-8
# an arc prices out when its reduced cost is >= -this
TOLERANCE_FOR_REDUCED_COSTS = 1e-6
# column generation multiplies the restricted-master size by this every round
COLUMN_GENERATION_RATIO = 2
