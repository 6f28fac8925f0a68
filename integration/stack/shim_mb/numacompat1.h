/* numacompat1.h -- see numa.h (same stand-in) */
#include "numa.h"
