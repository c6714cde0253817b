// opencv2/core/eigen.hpp -- empty: parameters.h includes it, the factor sources use nothing from it.
