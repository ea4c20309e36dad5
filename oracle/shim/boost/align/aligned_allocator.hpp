// Stand-in for boost::alignment::aligned_allocator (oracle build only).
#ifndef B200_ORACLE_BOOST_ALIGNED_ALLOCATOR
#define B200_ORACLE_BOOST_ALIGNED_ALLOCATOR
#include <cstddef>
#include <new>
#include <boost/align/aligned_alloc.hpp>
namespace boost { namespace alignment {
template <class T, std::size_t Alignment = 64> class aligned_allocator {
public:
	typedef T value_type; typedef T* pointer; typedef const T* const_pointer;
	typedef T& reference; typedef const T& const_reference;
	typedef std::size_t size_type; typedef std::ptrdiff_t difference_type;
	template <class U> struct rebind { typedef aligned_allocator<U, Alignment> other; };
	aligned_allocator() noexcept {}
	template <class U> aligned_allocator(const aligned_allocator<U,Alignment>&) noexcept {}
	pointer allocate(size_type n, const void* = 0) {
		void *p = boost::alignment::aligned_alloc(Alignment, n*sizeof(T));
		if(!p) throw std::bad_alloc();
		return static_cast<T*>(p);
	}
	void deallocate(pointer p, size_type) { boost::alignment::aligned_free(p); }
	template <class U, class... Args> void construct(U *p, Args&&... args) { ::new((void*)p) U(static_cast<Args&&>(args)...); }
	template <class U> void destroy(U *p) { p->~U(); }
	size_type max_size() const noexcept { return static_cast<size_type>(-1)/sizeof(T); }
};
template <class T, class U, std::size_t A>
inline bool operator==(const aligned_allocator<T,A>&, const aligned_allocator<U,A>&) noexcept { return true; }
template <class T, class U, std::size_t A>
inline bool operator!=(const aligned_allocator<T,A>&, const aligned_allocator<U,A>&) noexcept { return false; }
}}
#endif
