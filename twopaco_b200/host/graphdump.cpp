// graphdump.cpp -- the `graphdump` command line (B200 build): converts de_bruijn.bin to text on the GPU.  Same flags and
// messages as the reference's src/graphdump/graphdump.cpp:608-710 (TCLAP there; a small hand-rolled parser here):
//   -f/--format {seq|group|dot|gfa1|gfa2|fasta} (required)   -k/--kvalue <int> (required)
//   -s/--seqfile <fasta> (repeatable; required for gfa1 / gfa2 / fasta)   --prefix   <infile>
// The text goes to stdout; errors print "error: ..." to stderr and exit 1 (:696-705).
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/twopaco_b200.h"

namespace
{
	struct ArgError : std::runtime_error
	{
		ArgError(const std::string & msg, const std::string & arg) : std::runtime_error(msg), argId(arg) {}
		std::string argId;
	};

	void Usage(std::ostream & os)
	{
		os << "USAGE:\n   graphdump  -k <integer> [-s <string>] ... -f <seq|group|dot|gfa1|gfa2|fasta> [--prefix] <file name>\n\n"
			"   This utility converts the binary output of TwoPaCo to another format (B200 build)\n";
	}
}

int main(int argc, char * argv[])
{
	try
	{
		std::string format, inFile;
		std::vector<std::string> seqFiles;
		bool prefix = false, formatSet = false, kSet = false, inSet = false;
		unsigned k = 25;
		for (int i = 1; i < argc; ++i)
		{
			std::string a = argv[i], key, value;
			bool hasValue = false;
			if (a.size() < 2 || a[0] != '-')
			{
				if (inSet) throw ArgError("Too many arguments!", a);
				inFile = a; inSet = true;
				continue;
			}
			if (a[1] == '-')
			{
				size_t eq = a.find('=');
				std::string name = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
				key = name == "format" ? "f" : name == "seqfile" ? "s" : name == "kvalue" ? "k" : name == "prefix" ? "prefix" : name == "help" ? "h" : "";
				if (key.empty()) throw ArgError("Couldn't find match for argument", a);
				if (eq != std::string::npos) { value = a.substr(eq + 1); hasValue = true; }
			}
			else
			{
				key = a.substr(1, 1);
				if (std::string("fskh").find(key) == std::string::npos) throw ArgError("Couldn't find match for argument", a);
				if (a.size() > 2) { value = a.substr(2); hasValue = true; }
			}
			if (key == "prefix") { prefix = true; continue; }
			if (key == "h") { Usage(std::cout); return 0; }
			if (!hasValue)
			{
				if (i + 1 >= argc) throw ArgError("Missing a value for this argument!", a);
				value = argv[++i];
			}
			if (key == "f")
			{
				static const char * kFormats[] = { "seq", "group", "dot", "gfa1", "gfa2", "fasta" };
				bool ok = false;
				for (const char * f : kFormats) ok = ok || value == f;
				if (!ok) throw ArgError("Value '" + value + "' does not meet constraint: seq|group|dot|gfa1|gfa2|fasta", a);
				format = value; formatSet = true;
			}
			else if (key == "s") seqFiles.push_back(value);
			else if (key == "k")
			{
				std::istringstream ss(value);
				if (value.empty() || value[0] == '-' || !(ss >> k) || !ss.eof()) throw ArgError("Couldn't read argument value from string '" + value + "'", a);
				kSet = true;
			}
		}
		if (!inSet) throw ArgError("Required argument missing: infile", "infile");
		if (!formatSet) throw ArgError("Required argument missing: format", "-f (--format)");
		if (!kSet) throw ArgError("Required argument missing: kvalue", "-k (--kvalue)");

		int rc;
		if (format == "seq" || format == "group" || format == "dot")
		{
			rc = tpc_graphdump_file(inFile.c_str(), format.c_str(), "-");
		}
		else
		{
			if (seqFiles.empty()) throw ArgError("Required argument missing\n", "seqfilename");   // graphdump.cpp:668-671
			std::vector<const char *> paths;
			for (const std::string & s : seqFiles) paths.push_back(s.c_str());
			rc = tpc_graphdump_gfa_file(inFile.c_str(), format.c_str(), k, paths.data(), paths.size(), prefix ? 1 : 0, "-");
		}
		if (rc != 0) throw std::runtime_error(tpc_last_error());
	}
	catch (ArgError & e)
	{
		std::cerr << "error: " << e.what() << " for arg " << e.argId << std::endl;
		return 1;
	}
	catch (std::runtime_error & e)
	{
		std::cerr << "error: " << e.what() << std::endl;
		return 1;
	}

	return 0;
}
