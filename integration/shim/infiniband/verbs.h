#include "../../../oracle/shim/infiniband/verbs.h"  // the loopback fake verbs declarations (only needed because ev_posix.h drags verbs.h into most of core)
