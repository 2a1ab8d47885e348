// junctionapi.h -- reader / writer of the de_bruijn.bin junction-position stream, API-compatible
// with the reference's src/common/junctionapi.h (same class and method names) so that code
// written against it (graphdump.cpp:120-168, test.cpp:226-227) compiles against this header.
//
// Format (junctionapi.h:107-137 of the reference):
//   headerless stream of 12-byte little-endian units {u32 pos; i64 id};
//   a unit with pos == 0xFFFFFFFF or id == INT64_MAX is a separator and advances the sequence
//   index by one; before the first record of sequence c one separator is written for every
//   sequence index skipped since the last record; there are no trailing separators.
#ifndef TWOPACO_B200_JUNCTION_API_H_
#define TWOPACO_B200_JUNCTION_API_H_

#include <cstdint>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace TwoPaCo
{
	struct JunctionPosition
	{
		JunctionPosition() : chr_(UINT32_MAX), pos_(UINT32_MAX), bifId_(INT64_MAX) {}
		JunctionPosition(uint32_t chr, uint32_t pos, int64_t bifId) : chr_(chr), pos_(pos), bifId_(bifId) {}
		uint32_t GetPos() const { return pos_; }
		uint32_t GetChr() const { return chr_; }
		int64_t GetId() const { return bifId_; }

		static const size_t UNIT_BYTES = 12;
		static bool IsSeparator(uint32_t pos, int64_t id) { return pos == UINT32_MAX || id == INT64_MAX; }

	private:
		uint32_t chr_;
		uint32_t pos_;
		int64_t bifId_;
		friend class JunctionPositionReader;
		friend class JunctionPositionWriter;
	};

	class JunctionPositionReader
	{
	public:
		explicit JunctionPositionReader(const std::string & inFileName) : chr_(0), in_(inFileName.c_str(), std::ios::binary)
		{
			if (!in_)
			{
				throw std::runtime_error("Can't read the input file");
			}
		}

		// next non-separator unit; false at end of stream
		bool NextJunctionPosition(JunctionPosition & out)
		{
			char unit[JunctionPosition::UNIT_BYTES];
			while (in_.read(unit, sizeof(unit)))
			{
				uint32_t pos;
				int64_t id;
				std::memcpy(&pos, unit, 4);
				std::memcpy(&id, unit + 4, 8);
				if (JunctionPosition::IsSeparator(pos, id))
				{
					++chr_;
					continue;
				}

				out = JunctionPosition(chr_, pos, id);
				return true;
			}

			return false;
		}

		// mark[chr][pos] = true for every record of the stream
		void RestoreAllVectors(std::vector<std::vector<bool> > & mark)
		{
			JunctionPosition jp;
			while (NextJunctionPosition(jp))
			{
				mark.at(jp.GetChr()).at(jp.GetPos()) = true;
			}
		}

	private:
		uint32_t chr_;
		std::ifstream in_;
	};

	class JunctionPositionWriter
	{
	public:
		explicit JunctionPositionWriter(const std::string & outFileName) : chr_(0), out_(outFileName.c_str(), std::ios::binary)
		{
			if (!out_)
			{
				throw std::runtime_error("Can't create the output file");
			}
		}

		void WriteJunction(JunctionPosition jp)
		{
			while (chr_ < jp.chr_)
			{
				Put(UINT32_MAX, INT64_MAX);
				++chr_;
			}

			Put(jp.pos_, jp.bifId_);
			if (!out_)
			{
				throw std::runtime_error("Can't write to the output file");
			}
		}

	private:
		void Put(uint32_t pos, int64_t id)
		{
			char unit[JunctionPosition::UNIT_BYTES];
			std::memcpy(unit, &pos, 4);
			std::memcpy(unit + 4, &id, 8);
			out_.write(unit, sizeof(unit));
		}

		uint32_t chr_;
		std::ofstream out_;
	};
}

#endif
