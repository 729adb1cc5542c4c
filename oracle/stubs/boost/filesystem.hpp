// Build-harness stub (oracle/_ref only): the reference includes <boost/filesystem.hpp>
// (utils.cu:1) but boost headers are not installed in this image.  Every use in the
// reference (path, operator/, exists, create_directories, directory_iterator,
// is_regular_file, extension, filename) exists with the same spelling in std::filesystem.
#pragma once
#include <filesystem>
#include <iostream>
#include <string>
#include <vector>
namespace boost { namespace filesystem = std::filesystem; }
