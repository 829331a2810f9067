#include "KaminoTimer.h"
