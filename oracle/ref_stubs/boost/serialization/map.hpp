// Stand-in for <boost/serialization/map.hpp>, TEST INFRASTRUCTURE ONLY.
#pragma once
