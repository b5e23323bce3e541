/* stand-in for <opencv2/calib3d/calib3d.hpp>: see minicv.hpp (oracle test infrastructure) */
#include "../../minicv.hpp"
