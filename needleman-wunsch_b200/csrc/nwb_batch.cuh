/* nwb_batch.cuh -- batch kernel (placeholder until implemented). */
#pragma once
#include "nwb_device.cuh"
