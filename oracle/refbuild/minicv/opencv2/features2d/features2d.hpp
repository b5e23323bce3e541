/* stand-in for <opencv2/features2d/features2d.hpp>: see minicv.hpp (oracle test infrastructure) */
#include "../../minicv.hpp"
