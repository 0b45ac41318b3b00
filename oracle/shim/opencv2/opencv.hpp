// forwards to the oracle OpenCV shim (test infrastructure only)
#include "cvshim_more.h"
