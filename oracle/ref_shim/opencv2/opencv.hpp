// opencv2/opencv.hpp -- empty: parameters.h includes it, the factor sources use nothing from it.
