// Umbrella header of the B200 drop-in (same role as /root/reference/include/Jet.hpp).
#pragma once
#include "jet/Abort.hpp"
#include "jet/PathInfo.hpp"
#include "jet/SlicedContractor.hpp"
#include "jet/TaskBasedContractor.hpp"
#include "jet/Tensor.hpp"
#include "jet/TensorHelpers.hpp"
#include "jet/TensorNetwork.hpp"
#include "jet/TensorNetworkIO.hpp"
#include "jet/Utilities.hpp"
#include "jet/Version.hpp"
#include "jet/permute/Permuter.hpp"
