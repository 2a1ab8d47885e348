// vertexenumerator.cpp -- CreateEnumerator over the C ABI (replaces the reference's factory,
// src/graphconstructor/vertexenumerator.cpp:73-94, and VertexEnumeratorImpl's constructor).
#include "vertexenumerator.h"

namespace TwoPaCo
{
	namespace
	{
		void LogThunk(void * ctx, const char * text)
		{
			*static_cast<std::ostream*>(ctx) << text;
		}

		class GpuVertexEnumerator : public VertexEnumerator
		{
		public:
			explicit GpuVertexEnumerator(tpc_handle * handle) : handle_(handle) {}
			~GpuVertexEnumerator() { tpc_free(handle_); }
			size_t GetVerticesCount() const { return tpc_vertices(handle_); }
			int64_t GetId(const std::string & vertex) const { return tpc_get_id(handle_, vertex.c_str()); }
		private:
			GpuVertexEnumerator(const GpuVertexEnumerator &);
			void operator = (const GpuVertexEnumerator &);
			tpc_handle * handle_;
		};
	}

	std::unique_ptr<VertexEnumerator> CreateEnumerator(const std::vector<std::string> & fileName,
		size_t vertexLength,
		size_t filterSize,
		size_t hashFunctions,
		size_t rounds,
		size_t threads,
		size_t abundance,
		const std::string & tmpFileName,
		const std::string & outFileName,
		std::ostream & logStream)
	{
		std::vector<const char*> path;
		for (const std::string & fn : fileName)
		{
			path.push_back(fn.c_str());
		}

		tpc_handle * handle = 0;
		int rc = tpc_build(path.data(), path.size(), uint32_t(vertexLength), uint32_t(filterSize), uint32_t(hashFunctions),
			uint32_t(rounds), uint32_t(threads), uint64_t(abundance), tmpFileName.c_str(), outFileName.c_str(),
			LogThunk, &logStream, &handle);
		if (rc != 0)
		{
			throw std::runtime_error(tpc_last_error());
		}

		return std::unique_ptr<VertexEnumerator>(new GpuVertexEnumerator(handle));
	}
}
