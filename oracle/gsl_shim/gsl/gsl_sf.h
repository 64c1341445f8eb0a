/* GSL API shim (oracle/gsl_shim/gsl/gsl_shim.h): TEST INFRASTRUCTURE, see that file */
#include "gsl_shim.h"
