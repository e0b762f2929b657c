// empty shim: the reference includes this header but uses nothing from it
#include <boost/graph/adjacency_list.hpp>
