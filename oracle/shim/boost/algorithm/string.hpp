// Empty stand-in: the reference includes <boost/algorithm/string.hpp> (contig.h:25) but uses nothing from it.
#pragma once
