#include <boost/align/aligned_alloc.hpp>
#include <boost/align/aligned_allocator.hpp>
