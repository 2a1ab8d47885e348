// selftest.cpp -- see selftest.h.  The brute-force finder works from the definition of a
// junction: a k-mer over ACGT that, counting both strands, is preceded
// or followed by two different symbols, where every 'N' and every sequence end is a symbol of
// its own.
#include <algorithm>
#include <fstream>
#include <iostream>
#include <map>
#include <random>
#include <set>
#include <sstream>
#include <vector>

#include "junctionapi.h"
#include "selftest.h"
#include "vertexenumerator.h"

namespace TwoPaCo
{
	namespace
	{
		const char BASES[] = "ACGT";

		char Complement(char ch)
		{
			switch (ch)
			{
			case 'A': return 'T';
			case 'C': return 'G';
			case 'G': return 'C';
			case 'T': return 'A';
			}
			return 'N';
		}

		std::string ReverseComplement(const std::string & s)
		{
			std::string r(s.rbegin(), s.rend());
			std::transform(r.begin(), r.end(), r.begin(), Complement);
			return r;
		}

		bool Definite(const std::string & s)
		{
			return s.find_first_not_of(BASES) == std::string::npos;
		}

		// neighbour symbol sets of one canonical k-mer: 4 bits of bases + a count of unique symbols
		struct Neighbours
		{
			unsigned inBases, outBases, inUnique, outUnique;
			Neighbours() : inBases(0), outBases(0), inUnique(0), outUnique(0) {}
			static unsigned Bits(unsigned m) { unsigned c = 0; for (; m; m &= m - 1) ++c; return c; }
			bool Junction() const { return Bits(inBases) + inUnique > 1 || Bits(outBases) + outUnique > 1; }
		};

		void Note(unsigned & bases, unsigned & unique, char ch)
		{
			const char * p = std::char_traits<char>::find(BASES, 4, ch);
			if (p) bases |= 1u << (p - BASES); else ++unique;
		}

		void BruteForce(const std::vector<std::string> & chr, size_t k, std::set<std::string> & junction,
			std::vector<std::vector<bool> > & mark)
		{
			std::map<std::string, Neighbours> vertex;
			for (const std::string & s : chr)
			{
				for (size_t i = 0; i + k <= s.size(); ++i)
				{
					std::string v = s.substr(i, k);
					if (!Definite(v)) continue;
					char prev = i ? s[i - 1] : '$', next = i + k < s.size() ? s[i + k] : '$';
					std::string rc = ReverseComplement(v);
					if (v < rc)
					{
						Neighbours & n = vertex[v];
						Note(n.inBases, n.inUnique, prev);
						Note(n.outBases, n.outUnique, next);
					}
					else
					{
						Neighbours & n = vertex[rc];
						Note(n.inBases, n.inUnique, Complement(next));
						Note(n.outBases, n.outUnique, Complement(prev));
					}
				}
			}

			for (const auto & kv : vertex)
			{
				if (kv.second.Junction())
				{
					junction.insert(kv.first);
					junction.insert(ReverseComplement(kv.first));
				}
			}

			mark.assign(chr.size(), std::vector<bool>());
			for (size_t c = 0; c < chr.size(); ++c)
			{
				const std::string & s = chr[c];
				mark[c].assign(s.size(), false);
				for (size_t i = 0; i + k <= s.size(); ++i)
				{
					if (i == 0 || i + k == s.size() || junction.count(s.substr(i, k))) mark[c][i] = true;
				}
			}
		}
	}

	bool RunTests(size_t tests, size_t filterBits, size_t length, size_t chrNumber, const std::string & temporaryDir)
	{
		const std::string fasta = temporaryDir + "/test.fa";
		const std::string outBin = temporaryDir + "/out.bin";
		std::random_device rd;
		std::mt19937_64 rng(rd());
		std::uniform_real_distribution<> coin(0, 1);
		for (size_t t = 0; t < tests; ++t)
		{
			std::vector<std::string> chr(chrNumber);
			for (size_t i = 0; i < length; ++i) chr[0].push_back(rng() % 500 == 0 ? 'N' : BASES[rng() % 4]);
			for (size_t c = 1; c < chrNumber; ++c)
			{
				for (char ch : chr[0])
				{
					if (coin(rng) > 0.05) chr[c].push_back(ch);
					else if (coin(rng) <= 0.1) chr[c].push_back(BASES[rng() % 4]);
					else if (coin(rng) <= 0.5) { chr[c].push_back(ch); chr[c].push_back(BASES[rng() % 4]); }
				}
			}

			{
				std::ofstream f(fasta.c_str());
				if (!f) throw std::runtime_error("Can't create a temporary file for testing");
				for (size_t c = 0; c < chrNumber; ++c) f << ">" << c << "\n" << chr[c] << "\n";
			}

			for (size_t k = 3; k < 11; k += 2)
			{
				std::set<std::string> junction;
				std::vector<std::vector<bool> > naive;
				BruteForce(chr, k, junction, naive);
				for (size_t rounds = 1; rounds < 5; ++rounds)
				{
					std::stringstream null;
					std::unique_ptr<VertexEnumerator> vid = CreateEnumerator(std::vector<std::string>(1, fasta), k, filterBits, 1,
						rounds, 4, UINT64_MAX, temporaryDir, outBin, null);
					std::vector<std::vector<bool> > fast(chrNumber);
					for (size_t c = 0; c < chrNumber; ++c) fast[c].assign(chr[c].size(), false);
					JunctionPositionReader(outBin).RestoreAllVectors(fast);
					bool ok = fast == naive;
					for (const std::string & v : junction) ok = ok && vid->GetId(v) != INVALID_VERTEX;
					if (!ok)
					{
						std::cerr << "Test # " << t << " FAILED (k=" << k << ", rounds=" << rounds << ")" << std::endl;
						return false;
					}
				}
			}

			std::remove(fasta.c_str());
			std::remove(outBin.c_str());
			std::cerr << "Test # " << t << " PASSED" << std::endl;
		}

		return true;
	}
}
