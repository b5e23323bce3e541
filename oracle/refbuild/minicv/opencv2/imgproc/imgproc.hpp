/* stand-in for <opencv2/imgproc/imgproc.hpp>: see minicv.hpp (oracle test infrastructure) */
#include "../../minicv.hpp"
