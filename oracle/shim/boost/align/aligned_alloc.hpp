// Stand-in for boost::alignment::aligned_alloc/aligned_free (oracle build only).
#ifndef B200_ORACLE_BOOST_ALIGNED_ALLOC
#define B200_ORACLE_BOOST_ALIGNED_ALLOC
#include <cstdlib>
#include <cstddef>
namespace boost { namespace alignment {
inline void *aligned_alloc(std::size_t alignment, std::size_t size) {
	void *p = nullptr;
	if(alignment < sizeof(void*)) alignment = sizeof(void*);
	if(posix_memalign(&p, alignment, size ? size : alignment)) return nullptr;
	return p;
}
inline void aligned_free(void *p) { free(p); }
}}
#endif
