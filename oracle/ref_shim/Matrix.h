/* Case alias: the reference includes "Matrix.h" (M/LeastSquare.h:6) but the file is matrix.h. */
#include "matrix.h"
