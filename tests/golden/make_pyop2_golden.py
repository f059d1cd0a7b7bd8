"""Writes tests/golden/pyop2_test_matrices.json: the input data and golden
output arrays of the reference's PyOP2 value-level test
tests/pyop2/test_matrices.py (mesh :48-61, coords :108-113, f :123-125,
quadrature table of the C-string kernels :166-330, expected arrays :462-503).
These are DATA (a handful of numbers), transcribed; the kernels themselves are
restated in oracle/oracle.c.  The reference cannot be imported in this image
(PETSc/loopy absent), so this script only records the transcription."""
import json
import os

d = {
    "source": "reference tests/pyop2/test_matrices.py",
    "num_nodes": 4,
    "elem_node_map": [[0, 1, 3], [2, 3, 1]],
    "coords": [[0.0, 0.0], [2.0, 0.0], [1.0, 1.0], [0.0, 1.5]],
    "f": [1.0, 2.0, 3.0, 4.0],
    # 6-point rule, 8-digit table exactly as hard-coded in the reference kernels;
    # basis order {x, y, 1-x-y}
    "CG1": [[0.09157621, 0.09157621, 0.81684757, 0.44594849, 0.44594849, 0.10810302],
            [0.09157621, 0.81684757, 0.09157621, 0.44594849, 0.10810302, 0.44594849],
            [0.81684757, 0.09157621, 0.09157621, 0.10810302, 0.44594849, 0.44594849]],
    "d_CG1": [[1.0, 0.0], [0.0, 1.0], [-1.0, -1.0]],
    "w": [0.05497587, 0.05497587, 0.05497587, 0.11169079, 0.11169079, 0.11169079],
    "expected_matrix": [[0.25, 0.125, 0.0, 0.125],
                        [0.125, 0.291667, 0.0208333, 0.145833],
                        [0.0, 0.0208333, 0.0416667, 0.0208333],
                        [0.125, 0.145833, 0.0208333, 0.291667]],
    "expected_matrix_eps": 1e-5,
    "expected_rhs": [0.9999999523522115, 1.3541666031724144, 0.2499999883507239,
                     1.6458332580869566],
    "expected_rhs_eps": 1e-12,
}
with open(os.path.join(os.path.dirname(__file__), "pyop2_test_matrices.json"), "w") as fh:
    json.dump(d, fh, indent=1)
