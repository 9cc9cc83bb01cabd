// std::filesystem stand-in for the boost::filesystem names the reference headers use (TEST INFRASTRUCTURE).
#ifndef APDE_STUB_BOOST_FS_HPP_
#define APDE_STUB_BOOST_FS_HPP_
#include <filesystem>
#include <fstream>
namespace boost { namespace filesystem {
using path = std::filesystem::path;
using std::filesystem::exists;
using std::filesystem::create_directory;
using std::ifstream;
using std::ofstream;
} }
#endif
