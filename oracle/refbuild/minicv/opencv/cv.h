/* stand-in for <opencv/cv.h>: see minicv.hpp (oracle test infrastructure) */
#include "../minicv.hpp"
