"""Orchestration glue on the hot path: the `oisatgmi` class of
/root/reference/oisatgmi/driver.py:17-114 with the same method names and
attributes (`recal_amf`, `conv_ak`, `average`, `bias_correct`, `oi`), wired to
the GPU drop-ins.  File I/O (`read_data`, driver.py:22-34), reporting and NetCDF
output are out of scope: `read_data`
delegates to the reference's own reader when that package is importable, and
`attach` accepts an in-memory reader object (anything with `.ctm_data` and
`.sat_data`), which is what the tests and bench.py use.
"""
from __future__ import annotations

from .ak_conv_gosat import ak_conv_gosat
from .ak_conv_mopitt import ak_conv_mopitt
from .amf_recal import amf_recal
from .averaging import averaging
from .optimal_interpolation import OI

# validation-based affine corrections y -> (y - a) / b, driver.py:65-106
BIAS_CORRECTION = {
    ("TROPOMI", "NO2"): (0.32, 0.66),
    ("TROPOMI", "HCHO"): (0.90, 0.59),
    ("OMI", "NO2"): (0.32, 0.63),
    ("OMI", "HCHO"): (0.821, 0.79),
}


class oisatgmi(object):

    def __init__(self) -> None:
        pass

    def attach(self, reader_obj, gasname: str):
        self.reader_obj = reader_obj
        self.gasname = gasname

    def read_data(self, ctm_type, ctm_path, ctm_gas_name, ctm_frequency, sat_type, sat_path,
                  YYYYMM, averaging=False, read_ak=True, trop=False, num_job=1, mcip_dir=None,
                  tempo_hour=None):
        try:
            from oisatgmi.reader import readers  # the reference's file readers
        except Exception as exc:  # netCDF4/h5py are not part of this package
            raise RuntimeError(
                "file readers are outside the accelerated hot path; install the reference "
                "package for read_data(), or hand in-memory data to attach()") from exc
        reader_obj = readers()
        reader_obj.add_ctm_data(ctm_type, ctm_path, mcip_dir=mcip_dir)
        reader_obj.read_ctm_data(YYYYMM, ctm_gas_name, frequency_opt=ctm_frequency,
                                 averaging=averaging, num_job=num_job)
        reader_obj.add_satellite_data(sat_type, sat_path)
        reader_obj.read_satellite_data(YYYYMM, read_ak=read_ak, trop=trop, num_job=num_job,
                                       tempo_hour=tempo_hour)
        self.reader_obj = reader_obj
        self.gasname = ctm_gas_name[0]

    def recal_amf(self):
        self.reader_obj.sat_data = amf_recal(self.reader_obj.ctm_data, self.reader_obj.sat_data)

    def cal_pwv(self):
        from .pwv_cal import pwv_calculator
        self.reader_obj.sat_data = pwv_calculator(self.reader_obj.ctm_data,
                                                  self.reader_obj.sat_data)

    def conv_ak(self, sensor: str):
        if sensor == 'MOPITT':
            self.reader_obj.sat_data = ak_conv_mopitt(self.reader_obj.ctm_data,
                                                      self.reader_obj.sat_data)
        if sensor == 'GOSAT':
            self.reader_obj.sat_data = ak_conv_gosat(self.reader_obj.ctm_data,
                                                     self.reader_obj.sat_data)

    def average(self, startdate: str, enddate: str, gasname=None):
        (self.sat_averaged_vcd, self.sat_averaged_error, self.ctm_averaged_vcd, self.aux1,
         self.aux2, self.avg_time) = averaging(startdate, enddate, self.reader_obj)
        if gasname == 'O3':
            self.ctm_averaged_vcd = self.ctm_averaged_vcd / (2.69e16 * 1e-15)  # driver.py:62-63

    def bias_correct(self, sat_type, gasname):
        if (sat_type, gasname) in BIAS_CORRECTION:
            a, b = BIAS_CORRECTION[(sat_type, gasname)]
            self.sat_averaged_vcd = (self.sat_averaged_vcd - a) / b

    def oi(self, sensor: str, error_ctm=50.0):
        if sensor != 'GOSAT':
            xa, y = self.ctm_averaged_vcd, self.sat_averaged_vcd
        else:
            xa, y = self.aux2, self.aux1  # driver.py:113-114
        (self.ctm_averaged_vcd_corrected, self.ak_OI, self.increment_OI, self.error_OI) = OI(
            xa, y, (xa * error_ctm / 100.0) ** 2, self.sat_averaged_error ** 2,
            regularization_on=True)

    OUTPUT_NAMES = ("sat_averaged_vcd", "ctm_averaged_vcd_prior", "ctm_averaged_vcd_posterior",
                    "sat_averaged_error", "ak_OI", "error_OI", "scaling_factor", "aux1", "aux2")

    def output_fields(self):
        """The float32 variables driver.write_to_nc stores (driver.py:180-222), made on
        the GPU in one pass (K8): name -> float32 array, plus lon/lat of the first
        granule and the time string, as the reference writes them."""
        import numpy as np
        from . import _dev, _lib
        _dev.require_cuda()
        shape = np.shape(self.sat_averaged_vcd)
        src = [self.sat_averaged_vcd, self.ctm_averaged_vcd, self.ctm_averaged_vcd_corrected,
               self.sat_averaged_error, self.ak_OI, self.error_OI, self.aux1, self.aux2]
        dev = [_dev.to_device(np.ascontiguousarray(a, dtype=np.float64).ravel()) for a in src]
        n = dev[0].numel()
        out = _dev.empty((9, n), "float32")
        _lib.check(_lib.lib().oisat_output_fields(n, *[d.data_ptr() for d in dev], out.data_ptr(),
                                                  _dev.stream()))
        host = _dev.to_host(out)
        fields = {name: host[k].reshape(shape) for k, name in enumerate(self.OUTPUT_NAMES)}
        first = next(s for s in self.reader_obj.sat_data if s is not None)
        fields["lon"] = np.asarray(first.longitude_center).astype(np.float32)
        fields["lat"] = np.asarray(first.latitude_center).astype(np.float32)
        fields["time"] = self.avg_time.strftime("%Y-%m-%d %H:%M:%S")
        return fields

    def write_to_nc(self, output_file, output_folder='diag'):
        """driver.py:156-227.  The NetCDF container is file I/O (netCDF4, the reference's
        dependency, outside this package): written when that module is importable."""
        import os
        import numpy as np
        try:
            from netCDF4 import Dataset
        except Exception as exc:
            raise RuntimeError("write_to_nc needs the netCDF4 package; output_fields() returns "
                               "the variables it would store") from exc
        fields = self.output_fields()
        os.makedirs(output_folder, exist_ok=True)
        nc = Dataset(output_folder + '/' + output_file + '.nc', 'w')
        nx, ny = np.shape(fields["sat_averaged_vcd"])
        nc.createDimension('x', nx)
        nc.createDimension('y', ny)
        nc.createDimension('t', None)
        nc.createVariable('time', 'S1', ('t'))[:] = np.array(list(fields["time"]), 'S1')
        order = self.OUTPUT_NAMES[:7] + ("lon", "lat") + self.OUTPUT_NAMES[7:]
        for name in order:
            nc.createVariable(name, 'f', ('x', 'y'))[:, :] = fields[name]
        nc.close()
