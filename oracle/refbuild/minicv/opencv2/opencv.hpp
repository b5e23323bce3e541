/* stand-in for <opencv2/opencv.hpp>: see minicv.hpp (oracle test infrastructure) */
#include "../minicv.hpp"
