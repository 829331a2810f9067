#include "Kamino.h"
