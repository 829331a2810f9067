#include "KaminoSolver.h"
