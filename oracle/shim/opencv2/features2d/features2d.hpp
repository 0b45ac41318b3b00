#include "cvshim.h"
