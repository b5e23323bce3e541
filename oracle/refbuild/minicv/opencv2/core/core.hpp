/* stand-in for <opencv2/core/core.hpp>: see minicv.hpp (oracle test infrastructure) */
#include "../../minicv.hpp"
