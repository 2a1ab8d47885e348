// constructor.cpp -- the `twopaco` command line (B200 build).  Same flags, defaults and exit
// codes as the reference's src/graphconstructor/constructor.cpp:53-218 (TCLAP there; a small
// hand-rolled parser here):
//   -k/--kvalue <odd int, 25>   -f/--filtersize <bits>  XOR  --filtermemory <GB>
//   -q/--hashfnumber <5>  -r/--rounds <1>  -t/--threads <1>  -a/--abundance <2^64-1>
//   --tmpdir <.>  -o/--outfile <de_bruijn.bin>  --test  <fasta files...>
// Errors print "Error: ..." to stderr and exit 1 (constructor.cpp:179-188).
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "selftest.h"
#include "vertexenumerator.h"

namespace
{
	struct ArgError : std::runtime_error
	{
		ArgError(const std::string & msg, const std::string & arg) : std::runtime_error(msg), argId(arg) {}
		std::string argId;
	};

	struct Options
	{
		unsigned k = 25, q = 5, rounds = 1, threads = 1, filterSize = 32;
		double filterMemory = 4;
		bool filterSizeSet = false, filterMemorySet = false, test = false;
		uint64_t abundance = UINT64_MAX;
		std::string tmpDir = ".", outFile = "de_bruijn.bin";
		std::vector<std::string> files;
	};

	template<class T> T Parse(const std::string & text, const std::string & arg)
	{
		std::istringstream ss(text);
		T v;
		if (text.empty() || text[0] == '-' || !(ss >> v) || !ss.eof())
		{
			throw ArgError("Couldn't read argument value from string '" + text + "'", arg);
		}
		return v;
	}

	void Usage(std::ostream & os)
	{
		os << "USAGE:\n   twopaco  {-f <integer>|--filtermemory <float>} [-o <file name>] [--test] [--tmpdir <directory name>]\n"
			"            [-a <integer>] [-t <integer>] [-r <integer>] [-q <integer>] [-k <oddc>] <fasta files with genomes> ...\n\n"
			"   Program for construction of the condensed de Bruijn graph from complete genomes (B200 build)\n";
	}

	Options ParseArgs(int argc, char * argv[])
	{
		// long name -> canonical short key
		const std::map<std::string, std::string> longName = {
			{ "kvalue", "k" }, { "filtersize", "f" }, { "filtermemory", "filtermemory" }, { "hashfnumber", "q" },
			{ "rounds", "r" }, { "threads", "t" }, { "abundance", "a" }, { "tmpdir", "tmpdir" }, { "outfile", "o" },
			{ "test", "test" }, { "help", "h" }, { "version", "version" } };
		Options o;
		bool rest = false;
		for (int i = 1; i < argc; ++i)
		{
			std::string a = argv[i];
			if (rest || a.size() < 2 || a[0] != '-')
			{
				o.files.push_back(a);
				continue;
			}
			if (a == "--") { rest = true; continue; }
			std::string key, value;
			bool hasValue = false;
			if (a[1] == '-')
			{
				size_t eq = a.find('=');
				std::string name = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
				auto it = longName.find(name);
				if (it == longName.end()) throw ArgError("Couldn't find match for argument", a);
				key = it->second;
				if (eq != std::string::npos) { value = a.substr(eq + 1); hasValue = true; }
			}
			else
			{
				key = a.substr(1, 1);
				if (std::string("kfqrtaoh").find(key) == std::string::npos) throw ArgError("Couldn't find match for argument", a);
				if (a.size() > 2) { value = a.substr(2); hasValue = true; }
			}
			if (key == "test") { o.test = true; continue; }
			if (key == "h") { Usage(std::cout); std::exit(0); }
			if (key == "version") { std::cout << "\ntwopaco  version: 1.1.0 (twopaco_b200)\n\n"; std::exit(0); }
			if (!hasValue)
			{
				if (i + 1 >= argc) throw ArgError("Missing a value for this argument!", a);
				value = argv[++i];
			}
			if (key == "k")
			{
				o.k = Parse<unsigned>(value, a);
				if (o.k % 2 != 1) throw ArgError("Value '" + value + "' does not meet constraint: value of K must be odd", a);
			}
			else if (key == "f") { o.filterSize = Parse<unsigned>(value, a); o.filterSizeSet = true; }
			else if (key == "filtermemory") { o.filterMemory = Parse<double>(value, a); o.filterMemorySet = true; }
			else if (key == "q") o.q = Parse<unsigned>(value, a);
			else if (key == "r") o.rounds = Parse<unsigned>(value, a);
			else if (key == "t") o.threads = Parse<unsigned>(value, a);
			else if (key == "a") o.abundance = Parse<uint64_t>(value, a);
			else if (key == "tmpdir") o.tmpDir = value;
			else if (key == "o") o.outFile = value;
		}
		// cmd.xorAdd(filterSize, filterMemory): exactly one of the two (constructor.cpp:142)
		if (o.filterSizeSet == o.filterMemorySet)
		{
			throw ArgError(o.filterSizeSet ? "Mutually exclusive argument already set!" : "One (and only one) of the arguments -f / --filtermemory is required",
				"-f (--filtersize)");
		}
		if (o.files.empty()) throw ArgError("Required argument missing: filenames", "filenames");
		return o;
	}
}

int main(int argc, char * argv[])
{
	try
	{
		Options o = ParseArgs(argc, argv);
		if (o.test)
		{
			// constructor.cpp:145-149: 10 cases, 20 filter bits, 9000 bp, 6 sequences
			return TwoPaCo::RunTests(10, 20, 9000, 6, o.tmpDir) ? 0 : 1;
		}

		int64_t filterBits = o.filterSizeSet ? int64_t(o.filterSize) : int64_t(std::log2(o.filterMemory * 8e+9));  // constructor.cpp:151-159
		std::unique_ptr<TwoPaCo::VertexEnumerator> vid = TwoPaCo::CreateEnumerator(o.files, o.k, size_t(filterBits), o.q, o.rounds,
			o.threads, o.abundance, o.tmpDir, o.outFile, std::cout);
		if (vid)
		{
			std::cout << "Distinct junctions = " << vid->GetVerticesCount() << std::endl;
			std::cout << std::endl;
		}
	}
	catch (ArgError & e)
	{
		std::cerr << std::endl << "Error: " << e.what() << " for arg " << e.argId << std::endl;
		return 1;
	}
	catch (std::runtime_error & e)
	{
		std::cerr << std::endl << "Error: " << e.what() << std::endl;
		return 1;
	}

	return 0;
}
