// STAND-IN for muu/thread_pool.h: for_range runs the worker over [begin, end) on std::threads (static interleave) and
// tells the RNG shim which pixel index a worker invocation belongs to.  TEST INFRASTRUCTURE.
#pragma once
#include "preprocessor.h"
#include <cstddef>
#include <thread>
#include <vector>
namespace muu
{
	namespace shim
	{
		void on_pixel_begin(unsigned pixel_index) noexcept;
		extern unsigned row_step, row_width; // bounded samples: only rows y % row_step == 0 are rendered
	}
	class thread_pool
	{
		unsigned workers_;

	  public:
		explicit thread_pool(unsigned workers = 0) noexcept
			: workers_{ workers ? workers : (std::thread::hardware_concurrency() ? std::thread::hardware_concurrency() : 1u) }
		{}
		template <typename T, typename Func>
		void for_range(T begin, T end, Func&& func)
		{
			std::vector<std::thread> threads;
			auto body = [&](unsigned tid)
			{
				for (T i = begin + static_cast<T>(tid); i < end; i += static_cast<T>(workers_))
				{
					if (shim::row_step > 1 && shim::row_width && ((i / shim::row_width) % shim::row_step) != 0)
						continue;
					shim::on_pixel_begin(static_cast<unsigned>(i));
					func(i);
				}
			};
			for (unsigned t = 1; t < workers_; t++) threads.emplace_back(body, t);
			body(0);
			for (auto& t : threads) t.join();
		}
		void wait() noexcept {}
	};
}
