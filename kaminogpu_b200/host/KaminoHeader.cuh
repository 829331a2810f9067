#include "KaminoHeader.h"
