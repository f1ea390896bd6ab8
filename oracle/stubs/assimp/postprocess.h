// Empty stand-in (see Importer.hpp).
#pragma once
