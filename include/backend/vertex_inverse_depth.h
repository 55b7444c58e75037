// forwarding header: the whole drop-in API lives in myslam_backend_b200.h
#pragma once
#include "myslam_backend_b200.h"
