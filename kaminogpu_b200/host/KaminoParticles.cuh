#include "KaminoParticles.h"
