from oracle.thirdparty.o3 import *  # noqa: F401,F403
from oracle.thirdparty.o3 import (Irrep, Irreps, SphericalHarmonics, TensorProduct, FullyConnectedTensorProduct,  # noqa: F401
                                  ElementwiseTensorProduct, Linear, wigner_3j, rand_matrix, spherical_harmonics)
