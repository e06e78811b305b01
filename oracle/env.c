/* filled in below */
#include "prt_oracle.h"
