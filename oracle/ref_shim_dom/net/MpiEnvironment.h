// Shadow of net/MpiEnvironment.h (oracle/_ref Domain build only): ranks are threads started by fakempi_run.
#pragma once
namespace hemelb::net { class MpiEnvironment {}; }
