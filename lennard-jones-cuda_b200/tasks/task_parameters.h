// Parameter files of the task drivers: "<name> <value>" per line, '#' starts a comment, `output_prefix` takes a
// string.  Same format, names and defaults as the reference's RunFluctuationsParameters / RunIsothermParameters
// (/root/reference/src/tasks/run-fluctuations/include/run-fluctuations-aux.h:26-121,
//  /root/reference/src/tasks/run-isotherm/include/run-isotherm-aux.h:26-118; samples in /root/reference/input/).
#ifndef LJMD_TASKS_TASK_PARAMETERS_H
#define LJMD_TASKS_TASK_PARAMETERS_H
#include <cmath>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <map>
#include <sstream>
#include <string>

namespace ljtasks {

struct TaskParameters {
  std::string title;          // headline of the parameter listing
  std::string output_prefix;  // file-name stem
  std::map<std::string, double> values;

  double operator[](const std::string& k) const {
    std::map<std::string, double>::const_iterator it = values.find(k);
    return it == values.end() ? 0. : it->second;
  }
  long integer(const std::string& k) const { return std::lround((*this)[k]); }

  static TaskParameters fluctuations() {
    TaskParameters p;
    p.title = "Lennard-Jones Molecular Dynamics fluctuations run parameter list:";
    p.output_prefix = "run";
    p.values["N"] = 400;            p.values["T*"] = 1.4;       p.values["u*"] = 1.708;  p.values["rho*"] = 0.60;
    p.values["teq"] = 10.;          p.values["tfin"] = 1000.;   p.values["dt*"] = 0.004; p.values["canonical"] = 1;
    p.values["subvolume_spacing"] = 0.05;                       p.values["useCUDA"] = 1;
    return p;
  }
  static TaskParameters isotherm() {
    TaskParameters p;
    p.title = "Lennard-Jones Molecular Dynamics isotherm run parameter list:";
    p.output_prefix = "isotherm.run";
    p.values["N"] = 400;            p.values["T*"] = 1.4;       p.values["rho*_min"] = 0.60; p.values["rho*_max"] = 0.60;
    p.values["drho*"] = 0.01;       p.values["teq"] = 10.;      p.values["tfin"] = 5000.;    p.values["dt*"] = 0.004;
    p.values["useCUDA"] = 1;
    return p;
  }

  bool read(const std::string& path) {
    std::ifstream in(path.c_str());
    if (!in.is_open()) {
      std::cout << "Cannot open parameters file!" << "\n";
      return false;
    }
    std::cout << "Reading input parameters from file " << path << "\n";
    std::string key;
    while (in >> key) {
      if (key.empty() || key[0] == '#') {   // comment: drop the rest of the line
        in.ignore(std::numeric_limits<std::streamsize>::max(), '\n');
        continue;
      }
      std::cout << "Reading input parameter " << key << " = ";
      if (key == "output_prefix") {
        in >> output_prefix;
        std::cout << output_prefix << std::endl;
      } else {
        double v = 0.;
        in >> v;
        std::cout << v << std::endl;
        values[key] = v;
      }
    }
    std::cout.flush();
    return true;
  }

  void print() const {
    std::cout << title << "\n";
    for (std::map<std::string, double>::const_iterator it = values.begin(); it != values.end(); ++it)
      std::cout << std::setw(25) << it->first << " = " << it->second << "\n";
    std::cout << std::setw(25) << "output_prefix" << " = " << output_prefix << "\n" << std::endl;
  }

  // "<prefix>.<dd-mm-YYYY-THH-MM-SS>.N<N>.Tst<T*>|.ust<u*>[.rhost<rho*>]", as GetFullPrefix builds it
  std::string stamped_prefix(bool with_ensemble_and_density) const {
    std::time_t now = std::time(NULL);
    char stamp[64];
    std::strftime(stamp, sizeof(stamp), "%d-%m-%Y-T%H-%M-%S", std::localtime(&now));
    std::ostringstream ss;
    ss << output_prefix << "." << stamp << ".N" << (*this)["N"];
    if (!with_ensemble_and_density || integer("canonical")) ss << ".Tst" << (*this)["T*"];
    else ss << ".ust" << (*this)["u*"];
    if (with_ensemble_and_density) ss << ".rhost" << (*this)["rho*"];
    return ss.str();
  }
};

}  // namespace ljtasks
#endif
