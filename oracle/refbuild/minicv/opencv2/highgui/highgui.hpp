/* stand-in for <opencv2/highgui/highgui.hpp>: see minicv.hpp (oracle test infrastructure) */
#include "../../minicv.hpp"
