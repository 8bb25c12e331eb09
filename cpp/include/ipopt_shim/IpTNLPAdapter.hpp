// shim: see IpTNLP.hpp
#include "IpTNLP.hpp"
