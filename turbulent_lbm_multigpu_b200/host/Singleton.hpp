/* Singleton.hpp -- process-wide instance holder (reference src/Singleton.hpp): Instance() creates on first use. */
#ifndef LBM_B200_HOST_SINGLETON_HPP
#define LBM_B200_HOST_SINGLETON_HPP

template <typename C>
class Singleton {
public:
	static C *Instance()
	{
		static C the_instance;          /* thread-safe initialisation since C++11 */
		return &the_instance;
	}

private:
	Singleton();
};

#endif
