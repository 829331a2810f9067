#include "KaminoQuantity.h"
